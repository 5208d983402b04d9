-- Eckert V: the mean of the sinusoidal and the plate carree,
--   x = lon (1 + cos lat) / 2,   y = lat.                 Forward map only.
onload = "f_contain"
lens_height = pi
lens_width = pi*2
max_vfov = 180
max_fov = 360

local function project(lat, lon) return lon * (1 + cos(lat))/2, lat end

function lens_forward(rx, ry, rz)
  return project(ray_to_latlon(rx, ry, rz))
end
