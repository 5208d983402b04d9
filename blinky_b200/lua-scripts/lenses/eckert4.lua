-- Eckert IV equal-area projection.
-- Newton iteration (fixed 20 rounds) for the parametric angle
local function solve(lat)
  local t = lat/2
  local dt = 0
  for i = 1, 20 do
    dt = -(t + sin(t)*cos(t) + 2*sin(t) - (2+pi*0.5)*sin(lat))/(2*cos(t)*(1+cos(t)))
    t = t+dt
  end
  return t
end

-- half-width of the map at height y; cached per scanline
local last_y, half_width
local function row_half_width(y, lat)
  if y ~= last_y then
    local t = solve(abs(lat))
    half_width = 2/sqrt(pi*(4+pi))*pi*(1+cos(t))
    last_y = y
  end
  return half_width
end

local t = solve(pi*0.5)
local top = 2*sqrt(pi/(4+pi))*sin(t)

function lens_inverse(x, y)
  local t = asin(y/2*sqrt((4+pi)/pi))
  local lat = asin((t+sin(t)*cos(t)+2*sin(t))/(2+pi*0.5))
  local lon = sqrt(pi*(4+pi))*x/(2*(1+cos(t)))
  if abs(y) > top or abs(x) > row_half_width(y, lat) then
    return nil
  end
  return latlon_to_ray(lat, lon)
end

function lens_forward(x, y, z)
  local lat, lon = ray_to_latlon(x, y, z)
  local t = solve(lat)
  return 2/sqrt(pi*(4+pi))*lon*(1+cos(t)), 2*sqrt(pi/(4+pi))*sin(t)
end

max_fov = 360
max_vfov = 180

t = solve(0)
lens_width = 2/sqrt(pi*(4+pi))*pi*(1+cos(t))*2
lens_height = 2*top
onload = "f_contain"
