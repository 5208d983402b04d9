-- Miller cylindrical: Mercator applied to 4/5 of the latitude and stretched back by 5/4,
-- which keeps the poles on the map:  y = 5/4 ln tan(pi/4 + 2/5 lat).
onload = "f_contain"
max_vfov = 180
max_fov = 360

local top = 1.25*log(tan(0.25*pi+0.4*pi*0.5))   -- y of the pole
lens_height = top*2
lens_width = 2*pi

function lens_inverse(x, y)
  local off_map = abs(y) > top or abs(x) > pi
  if off_map then return nil end
  return latlon_to_ray(5/4*atan(sinh(4/5*y)), x)
end

function lens_forward(rx, ry, rz)
  local lat, lon = ray_to_latlon(rx, ry, rz)
  return lon, 1.25*log(tan(0.25*pi+0.4*lat))
end
