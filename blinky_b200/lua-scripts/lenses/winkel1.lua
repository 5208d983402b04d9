-- Winkel I: the mean of the sinusoidal and an equirectangular map whose standard
-- parallel makes cos(lat1) = 2/pi.      x = lon (cos lat1 + cos lat) / 2,  y = lat
-- Forward map only.
onload = "f_contain"
lens_height = pi
lens_width = pi * (2/pi + 1)/2 * 2   -- the equator: x(lat 0, lon pi), both sides
max_vfov = 180
max_fov = 360

local function project(lat, lon) return lon * (2/pi + cos(lat))/2, lat end

function lens_forward(rx, ry, rz)
  return project(ray_to_latlon(rx, ry, rz))
end
