-- Equirectangular / plate carree: the screen IS the (longitude, latitude) chart,
-- 2 pi wide and pi tall.  The format panorama tools exchange.
onload = "f_contain"
lens_height = pi
lens_width = 2*pi
max_vfov = 180
max_fov = 360

function lens_inverse(x, y)
  local off_chart = abs(y) > pi/2 or abs(x) > pi
  if off_chart then return nil end
  return latlon_to_ray(y, x)
end

function lens_forward(rx, ry, rz)
  local lat, lon = ray_to_latlon(rx, ry, rz)
  return lon, lat
end
