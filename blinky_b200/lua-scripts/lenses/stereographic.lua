-- Stereographic fisheye: the sphere projected from its far pole onto the image plane.
--
-- r = tan(theta/2).  Conformal (small shapes keep their shape), and it never ends:
-- there is no rim, only the antipode at infinity, hence no lens_width/height and a
-- plain "f_fov 180" as the default zoom.
onload = "f_fov 180"
max_vfov = 360
max_fov = 360

local half = 0.5   -- the 1/2 of theta/2

function lens_inverse(x, y)
  local r = sqrt(x*x+y*y)
  local theta = atan(r)/half
  local s = sin(theta)
  return x/r*s, y/r*s, cos(theta)
end

function lens_forward(rx, ry, rz)
  local r = tan(acos(rz)*half)
  local k = r/sqrt(rx*rx+ry*ry)
  return rx*k, ry*k
end
