-- Rectilinear (ordinary perspective): r = tan(theta).
max_fov = 180
max_vfov = 180
onload = "f_fov 110"

function lens_inverse(x, y)
  local r = sqrt(x*x+y*y)
  local theta = atan(r)
  local s = sin(theta)
  return x/r*s, y/r*s, cos(theta)
end

function lens_forward(x, y, z)
  local theta = acos(z)
  local r = tan(theta)
  local c = r/sqrt(x*x+y*y)
  return x*c, y*c
end
