-- Sinusoidal (Sanson-Flamsteed) equal-area map: parallels keep their true length,
-- x = lon cos lat, y = lat.  Forward map only.
onload = "f_contain"
lens_height = pi
lens_width = 2*pi
max_vfov = 180
max_fov = 360

local function project(lat, lon) return lon*cos(lat), lat end

function lens_forward(rx, ry, rz)
  return project(ray_to_latlon(rx, ry, rz))
end
