-- Equisolid-angle fisheye (the classic "mirror ball" / most real fisheye glass).
--
-- r = 2 sin(theta/2): equal solid angles cover equal image areas.  The antipode
-- lands on the rim circle of radius 2 sin(pi/2).
onload = "f_contain"
max_vfov = 360
max_fov = 360

local rim = 2*sin(pi*0.5)
lens_height = rim*2
lens_width = rim*2

local function angle_of(r) return 2*asin(r*0.5) end
local function radius_of(theta) return 2*sin(theta*0.5) end

function lens_inverse(x, y)
  local r = sqrt(x*x+y*y)
  if r > rim then return nil end
  local theta = angle_of(r)
  local s = sin(theta)
  return x/r*s, y/r*s, cos(theta)
end

function lens_forward(rx, ry, rz)
  local k = radius_of(acos(rz))/sqrt(rx*rx+ry*ry)
  return rx*k, ry*k
end
