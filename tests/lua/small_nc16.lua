-- a small run for the CPU tests of the command line (tests/test_lua_front.py): nc = args[1] (default 16), force mesh (2 nc)^3, COLA,
-- args[2] steps (default 4), args[3] particle_fraction (default 1), args[4] sort_snapshot (default true)
nc = tonumber(args[1] or "16")
boxsize = 2 * nc

time_step = linspace(0.1, 1, tonumber(args[2] or "4"))
output_redshifts = {0.0}

Omega_m = 0.307494
h       = 0.6774

read_powerspectrum = "powerspec.txt"
random_seed = 7

force_mode = "cola"
growth_mode = "LCDM"
pm_nc_factor = 2
lpt_nc_factor = 1
np_alloc_factor = 3.0

write_snapshot = "out/fastpm"
write_powerspectrum = "out/powerspec"
particle_fraction = tonumber(args[3] or "1.0")     -- < 1: the snapshot keeps the particles whose rand deviate is below it
sort_snapshot = (args[4] or "true") == "true"
