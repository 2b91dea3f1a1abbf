/* Prints sizeof / offsetof of the public structs of the libfastpm API for the hot path.  Compiled twice by
 * tests/test_abi_layout.py: against the reference's own headers (-DUSE_REFERENCE, /root/reference/api with the oracle's
 * MPI shim) and against include/fastpm_b200_api.h; the two outputs must be identical. */
#include <stdio.h>
#include <stddef.h>
#ifdef USE_REFERENCE
#include <mpi.h>
#include <fastpm/libfastpm.h>
#include <fastpm/logging.h>
#include <fastpm/prof.h>
#else
#include "fastpm_b200_api.h"
#endif

#define S(T) printf("sizeof %s %zu\n", #T, sizeof(T))
#define O(T, f) printf("offsetof %s.%s %zu\n", #T, #f, offsetof(T, f))

int main(void)
{
    S(FastPMFloat); S(FastPMEvent); S(FastPMEventHandler); S(FastPMMemory); S(PMRegion);
    S(FastPMStore); S(FastPMFieldDescr); S(FastPMPainter); S(FastPMCosmology); S(FastPMGrowthInfo);
    S(VPMInit); S(FastPMConfig); S(FastPMSolver); S(FastPMKickFactor); S(FastPMDriftFactor);
    S(FastPMForceEvent); S(FastPMLPTEvent); S(FastPMTransitionEvent); S(FastPMInterpolationEvent);
    S(FastPMState); S(FastPMStates); S(FastPMTransition); S(FastPMPowerSpectrum); S(FastPMFuncK);
    O(FastPMStore, attributes); O(FastPMStore, _base); O(FastPMStore, np); O(FastPMStore, np_upper); O(FastPMStore, _column_info);
    O(FastPMStore, meta); O(FastPMStore, columns); O(FastPMStore, x); O(FastPMStore, v); O(FastPMStore, acc); O(FastPMStore, dx1);
    O(FastPMStore, dx2); O(FastPMStore, id); O(FastPMStore, potential); O(FastPMStore, mass); O(FastPMStore, mask);
    O(FastPMPainter, pm); O(FastPMPainter, paint); O(FastPMPainter, readout); O(FastPMPainter, support); O(FastPMPainter, shift);
    O(FastPMConfig, nc); O(FastPMConfig, boxsize); O(FastPMConfig, alloc_factor); O(FastPMConfig, lpt_nc_factor);
    O(FastPMConfig, cosmology); O(FastPMConfig, vpminit); O(FastPMConfig, USE_DX1_ONLY); O(FastPMConfig, USE_SHIFT);
    O(FastPMConfig, ExtraAttributes); O(FastPMConfig, nLPT); O(FastPMConfig, PAINTER_TYPE); O(FastPMConfig, painter_support);
    O(FastPMConfig, FORCE_TYPE); O(FastPMConfig, KERNEL_TYPE); O(FastPMConfig, SOFTENING_TYPE); O(FastPMConfig, pgdc);
    O(FastPMSolver, basepm); O(FastPMSolver, lptpm); O(FastPMSolver, comm); O(FastPMSolver, NTask); O(FastPMSolver, ThisTask);
    O(FastPMSolver, species); O(FastPMSolver, has_species); O(FastPMSolver, cdm); O(FastPMSolver, config); O(FastPMSolver, cosmology);
    O(FastPMSolver, event_handlers); O(FastPMSolver, vpm_list);
    O(FastPMKickFactor, forcemode); O(FastPMKickFactor, nsamples); O(FastPMKickFactor, dda); O(FastPMKickFactor, Dv1); O(FastPMKickFactor, q1);
    O(FastPMKickFactor, ai); O(FastPMDriftFactor, dyyy); O(FastPMDriftFactor, da1); O(FastPMDriftFactor, Dv1); O(FastPMDriftFactor, ai);
    O(FastPMForceEvent, kernel); O(FastPMForceEvent, painter); O(FastPMForceEvent, pm); O(FastPMForceEvent, delta_k); O(FastPMForceEvent, N);
    O(FastPMForceEvent, a_f); O(FastPMForceEvent, a_n);
    O(FastPMTransition, action); O(FastPMTransition, a); O(FastPMTransition, i); O(FastPMTransition, start); O(FastPMTransition, end);
    O(FastPMCosmology, h); O(FastPMCosmology, Omega_m); O(FastPMCosmology, Omega_Lambda); O(FastPMCosmology, T_cmb); O(FastPMCosmology, N_nu);
    O(FastPMCosmology, m_ncdm); O(FastPMCosmology, growth_mode); O(FastPMCosmology, FDinterp);
    O(FastPMPowerSpectrum, base); O(FastPMPowerSpectrum, edges); O(FastPMPowerSpectrum, pm); O(FastPMPowerSpectrum, k0); O(FastPMPowerSpectrum, Nmodes);
    printf("enum FASTPM_FORCE_COLA %d FASTPM_KERNEL_1_4 %d FASTPM_PAINTER_CIC %d COLUMN_ACC %ld COLUMN_MASS %ld\n",
           (int) FASTPM_FORCE_COLA, (int) FASTPM_KERNEL_1_4, (int) FASTPM_PAINTER_CIC, (long) COLUMN_ACC, (long) COLUMN_MASS);
    printf("event names %s %s %s %s\n", FASTPM_EVENT_FORCE, FASTPM_EVENT_LPT, FASTPM_EVENT_TRANSITION, FASTPM_EVENT_INTERPOLATION);
    return 0;
}
