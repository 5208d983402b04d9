-- Mercator: the conformal cylinder.  x = longitude, y = ln tan(pi/4 + lat/2), whose
-- inverse is the Gudermannian lat = atan(sinh y).  Endless towards the poles: only a
-- width is declared and the default zoom covers the screen.
onload = "f_cover"
lens_width = 2*pi
max_vfov = 180
max_fov = 360

local function gudermannian(y) return atan(sinh(y)) end

function lens_inverse(x, y)
  if abs(x) > pi then return nil end   -- one turn around the cylinder
  return latlon_to_ray(gudermannian(y), x)
end

function lens_forward(rx, ry, rz)
  local lat, lon = ray_to_latlon(rx, ry, rz)
  return lon, log(tan(pi*0.25+lat*0.5))
end
