/* fastpm_b200 -- <fastpm/libfastpm.h> as seen by the reference's generated parameter reader (lua-config.c names the enum
 * constants of the schema, src/lua-runtime-fastpm.lua): this build's mirror of the libfastpm API plus the two option enums
 * whose features are outside the force-step path (api/fastpm/pngaussian.h:3-6, thermalvelocity.h:1-4) -- the run loop refuses
 * every value but the first of each. */
#ifndef FASTPM_B200_LUA_LIBFASTPM_H
#define FASTPM_B200_LUA_LIBFASTPM_H
#include "fastpm_b200_api.h"
typedef enum { FASTPM_FNL_NONE, FASTPM_FNL_LOCAL } FastPMPNGaussianType;
typedef enum { FASTPM_NCDM_SPHERE_HEALPIX = 0, FASTPM_NCDM_SPHERE_FIBONACCI = 1 } FastPMncdmSphereScheme;
#endif
