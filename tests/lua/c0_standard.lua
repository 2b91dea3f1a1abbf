-- BASELINE.json configs[0]: the reference's tests/standard.lua reduced to nc = 64 / boxsize = 128 Mpc/h, 5 PM steps,
-- with the same argument convention (first argument: za | 2lpt | cola | pm | fastpm; further keywords select options).
nc = 64
boxsize = 128.0

if args[1] == 'za' then
    za = true
    force_mode = "pm"
    time_step = {1.0}
elseif args[1] == '2lpt' then
    za = false
    force_mode = "pm"
    time_step = {1.0}
elseif args[1] == 'cola' then
    za = false
    force_mode = "cola"
    time_step = linspace(0.1, 1, 5)
elseif args[1] == 'pm' then
    za = false
    force_mode = "pm"
    time_step = linspace(0.1, 1, 5)
elseif args[1] == 'fastpm' then
    za = false
    force_mode = "fastpm"
    time_step = linspace(0.1, 1, 5)
else
    error("wrong arg!")
end

local function has(keyword)
    for i,k in pairs(args) do
        if k == keyword then
            return true
        end
    end
    return false
end
if has('lanczos2') then
    painter_type = "lanczos"
    painter_support = 4
end
if has('remove_variance') then
    remove_cosmic_variance = true
end
if has('gaussian36') then
    dealiasing_type = 'gaussian36'
end
if has('lightcone') then
    lc_write_usmesh = "c0/lightcone"
end

prefix = 'c0'
for i,k in pairs(args) do
    if i > 0 then
        prefix = prefix .. '-' .. k
    end
end

output_redshifts = {1.0, 0.0}

Omega_m = 0.307494
h       = 0.6774
read_powerspectrum = "powerspec.txt"
random_seed = 100

pm_nc_factor = {{0.0, 2}, {0.5, 3}}     -- variable force mesh: 128^3 up to a = 0.5, then 192^3 (vpm.c)
lpt_nc_factor = 1
np_alloc_factor = 2.0

write_snapshot = prefix .. "/fastpm"
write_powerspectrum = prefix .. "/powerspec"
