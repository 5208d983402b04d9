-- Miller cylindrical: Mercator with latitude scaled by 4/5.
local top = 1.25*log(tan(0.25*pi+0.4*pi*0.5))

max_fov = 360
max_vfov = 180
lens_width = 2*pi
lens_height = top*2
onload = "f_contain"

function lens_inverse(x, y)
  if abs(y) > top or abs(x) > pi then
    return nil
  end
  local lon = x
  local lat = 5/4*atan(sinh(4/5*y))
  return latlon_to_ray(lat, lon)
end

function lens_forward(x, y, z)
  local lat, lon = ray_to_latlon(x, y, z)
  return lon, 1.25*log(tan(0.25*pi+0.4*lat))
end
