-- The configuration of the reference's tests/lightcone.lua (nc = 64, box 512 Mpc/h, one mesh cell per particle, FastPM steps,
-- LCDM growth, seed 100 with the cosmic variance removed, 8 steps linspace(0.1, 1, 8)) WITHOUT its light-cone and FOF outputs,
-- which are outside the force-step path.  The reference's golden log of that run (tests/run-test-lightcone.check) fixes the
-- lines this run must print: white-noise variance, dx1 / dx2, and D^2 P(k<...) at every step.
nc = 64
boxsize = 512

time_step = linspace(0.1, 1, 8)
output_redshifts = {0.0}
compute_potential = true

Omega_m = 0.307494
h       = 0.6774

read_powerspectrum = "powerspec.txt"
random_seed = 100
remove_cosmic_variance = true

force_mode = "fastpm"
growth_mode = "LCDM"
pm_nc_factor = 1
lpt_nc_factor = 1
np_alloc_factor = 2.0

write_snapshot = "golden_nc64/fastpm"
write_powerspectrum = "golden_nc64/powerspec"
particle_fraction = 1.0
