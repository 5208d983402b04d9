-- Hammer (Hammer-Aitoff) equal-area world map: the whole sphere in a 2:1 ellipse
-- with semi-axes 2 sqrt 2 and sqrt 2.
--
--   forward: d = sqrt(1 + cos lat cos(lon/2))
--            x = 2 sqrt2 cos lat sin(lon/2) / d,   y = sqrt2 sin lat / d
--   inverse: z = sqrt(1 - x^2/16 - y^2/4)
--            lon = 2 atan(z x / (2 (2 z^2 - 1))),   lat = asin(z y)
onload = "f_contain"
max_vfov = 180
max_fov = 360
lens_height = sqrt(2)*2
lens_width = 2*sqrt(2)*2

local function inside_ellipse(x, y) return not (x*x/8+y*y/2 > 1) end

function lens_inverse(x, y)
  if not inside_ellipse(x, y) then return nil end
  local z = sqrt(1-0.0625*x*x-0.25*y*y)
  local lat = asin(z*y)
  return latlon_to_ray(lat, 2*atan(z*x/(2*(2*z*z-1))))
end

local function project(lat, lon)
  return 2*sqrt(2)*cos(lat)*sin(lon*0.5) / sqrt(1+cos(lat)*cos(lon*0.5)),
         sqrt(2)*sin(lat) / sqrt(1+cos(lat)*cos(lon*0.5))
end

function lens_forward(rx, ry, rz)
  return project(ray_to_latlon(rx, ry, rz))
end
