-- Gumby: the Pannini projection with angles stretched by 4/3 so the whole sphere fits.
local d = 1
local stretch = 0.75
local unstretch = 1.0/stretch

max_fov = 360
max_vfov = 180
onload = "f_contain"

function lens_inverse(x, y)
  local k = x*x/((d+1)*(d+1))
  local dscr = k*k*d*d - (k+1)*(k*d*d-1)
  local clon = (-k*d+sqrt(dscr))/(k+1)
  local S = (d+1)/(d+clon)
  local lon = atan2(x,S*clon)
  local lat = atan2(y,S)
  lon = lon*unstretch
  lat = lat*unstretch
  return latlon_to_ray(lat, lon)
end

function lens_forward(x, y, z)
  local lat, lon = ray_to_latlon(x, y, z)
  lon = lon*stretch
  lat = lat*stretch
  local S = (d+1)/(d+cos(lon))
  return S*sin(lon), S*tan(lat)
end

-- extent of the image: project the pole and the antimeridian
local _, top = lens_forward(latlon_to_ray(pi/2, 0))
lens_height = top*2
local right = lens_forward(latlon_to_ray(0, pi))
lens_width = right*2
