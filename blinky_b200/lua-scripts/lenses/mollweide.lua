-- Mollweide (homolographic) equal-area world map: a 2:1 ellipse, semi-axes 2 sqrt2 and sqrt2.
--
--   forward: solve 2t + sin 2t = pi sin lat for t, then
--            x = 2 sqrt2 / pi * lon cos t,   y = sqrt2 sin t
--   inverse: t = asin(y / sqrt2),  lon = pi x / (2 sqrt2 cos t),  lat = asin((2t + sin 2t)/pi)
onload = "f_contain"
max_vfov = 180
max_fov = 360
lens_height = sqrt(2)*2
lens_width = 2*sqrt(2)*2

local root2 = sqrt(2)

-- Newton's method on u + sin u = pi sin lat (u = 2t); stops on the first step below 1e-3
local function half_aux_angle(lat)
  local u = lat
  local step
  repeat
    step = -(u + sin(u) - pi*sin(lat))/(1+cos(u))
    u = u+step
  until step < 0.001
  return u/2
end

function lens_inverse(x, y)
  if x*x/8 + y*y/2 > 1 then return nil end   -- outside the ellipse
  local t = asin(y/root2)
  local lat = asin((2*t+sin(2*t))/pi)
  return latlon_to_ray(lat, pi*x/(2*root2*cos(t)))
end

function lens_forward(rx, ry, rz)
  local lat, lon = ray_to_latlon(rx, ry, rz)
  local t = half_aux_angle(lat)
  return 2*sqrt(2)/pi*lon*cos(t), sqrt(2)*sin(t)
end
