-- Gall stereographic cylindrical projection.
local YF  = 1.70710678118654752440
local XF  = 0.70710678118654752440
local RYF = 0.58578643762690495119
local RXF = 1.41421356237309504880
local right = XF * pi
local top = YF * tan(0.5*pi/2)

max_fov = 360
max_vfov = 180
lens_width = right*2
lens_height = top*2
onload = "f_contain"

function lens_forward(x, y, z)
  if abs(x) > right or abs(y) > top then
    return nil
  end
  local lat, lon = ray_to_latlon(x, y, z)
  return XF * lon, YF * tan(0.5 * lat)
end

function lens_inverse(x, y)
  local lon = RXF * x
  local lat = 2 * atan(y * RYF)
  return latlon_to_ray(lat, lon)
end
