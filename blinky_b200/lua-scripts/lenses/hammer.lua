-- Hammer equal-area projection (2:1 ellipse).
max_fov = 360
max_vfov = 180
lens_width = 2*sqrt(2)*2
lens_height = sqrt(2)*2
onload = "f_contain"

function lens_inverse(x, y)
  if x*x/8+y*y/2 > 1 then
    return nil  -- outside the ellipse
  end
  local z = sqrt(1-0.0625*x*x-0.25*y*y)
  local lon = 2*atan(z*x/(2*(2*z*z-1)))
  local lat = asin(z*y)
  return latlon_to_ray(lat, lon)
end

function lens_forward(x, y, z)
  local lat, lon = ray_to_latlon(x, y, z)
  local px = 2*sqrt(2)*cos(lat)*sin(lon*0.5) / sqrt(1+cos(lat)*cos(lon*0.5))
  local py = sqrt(2)*sin(lat) / sqrt(1+cos(lat)*cos(lon*0.5))
  return px, py
end
