-- Pannini projection (distance parameter d = 1): straight verticals, up to 360 degrees wide.
local d = 1

max_fov = 360
max_vfov = 180
onload = "f_fov 180"

function lens_inverse(x, y)
  local k = x*x/((d+1)*(d+1))
  local dscr = k*k*d*d - (k+1)*(k*d*d-1)
  local clon = (-k*d+sqrt(dscr))/(k+1)
  local S = (d+1)/(d+clon)
  local lon = atan2(x,S*clon)
  local lat = atan2(y,S)
  return latlon_to_ray(lat, lon)
end

function lens_forward(x, y, z)
  local lat, lon = ray_to_latlon(x, y, z)
  local S = (d+1)/(d+cos(lon))
  return S*sin(lon), S*tan(lat)
end
