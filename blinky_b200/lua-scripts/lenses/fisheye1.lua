-- Equidistant ("f-theta") fisheye.
--
-- The image radius IS the angle off the view axis, r = theta, so circles of equal
-- angular size stay equal all the way to the rim; the whole sphere fits in a disc
-- of radius pi.
--
--   inverse: theta = r,                    ray = (x/r sin theta, y/r sin theta, cos theta)
--   forward: theta = acos(z),  r = theta,  (x, y) scaled to length r
onload = "f_contain"
lens_height = 2*pi
lens_width = 2*pi
max_vfov = 360
max_fov = 360

-- unit ray for a point at radius `r`, direction (x, y), `theta` off axis
local function off_axis(x, y, r, theta)
  local s = sin(theta)
  return x/r*s, y/r*s, cos(theta)
end

function lens_inverse(x, y)
  local r = sqrt(x*x+y*y)
  if r > pi then return nil end   -- beyond the antipode
  return off_axis(x, y, r, r)
end

function lens_forward(rx, ry, rz)
  local r = acos(rz)
  local k = r/sqrt(rx*rx+ry*ry)
  return rx*k, ry*k
end
