-- Rectilinear: the ordinary pinhole camera, r = tan(theta).  Straight lines stay
-- straight, which is exactly why it cannot reach 180 degrees.
onload = "f_fov 110"
max_vfov = 180
max_fov = 180

function lens_inverse(x, y)
  local r = sqrt(x*x+y*y)
  local theta = atan(r)
  local s = sin(theta)
  return x/r*s, y/r*s, cos(theta)
end

function lens_forward(rx, ry, rz)
  local k = tan(acos(rz))/sqrt(rx*rx+ry*ry)
  return rx*k, ry*k
end
