-- Winkel tripel: forward in closed form, inverse by Newton iteration.
local clat0 = 2/pi  -- cosine of the standard parallel

max_fov = 360
max_vfov = 180
onload = "f_contain"

function lens_forward(x, y, z)
  local lat, lon = ray_to_latlon(x, y, z)
  local clat = cos(lat)
  local temp = clat*cos(lon*0.5)
  local D = acos(temp)
  local C = 1 - temp*temp
  temp = D/sqrt(C)
  local px = 0.5 * (2*temp*clat*sin(lon*0.5)+lon*clat0)
  local py = 0.5 * (temp*sin(lat) + lat)
  return px, py
end

local _, pole = lens_forward(latlon_to_ray(pi/2, 0))
lens_height = 2*pole
local edge = lens_forward(latlon_to_ray(0, pi))
lens_width = 2*edge

-- the iteration misbehaves in the far corners of the bounding box: cut them out
local cut_x = lens_width/2*0.71
local cut_y = lens_height/2*0.81

local eps = 0.0001
local halfpi = pi/2

-- Newton-Raphson on the forward equations in (lambda, phi), started at (x, y):
--   F(lambda, phi) = forward(lambda, phi) - (x, y) = 0
-- with the analytic Jacobian (Ipbuker & Bildirici, "A General Algorithm for the Inverse
-- Transformation of Map Projections Using Jacobian Matrices", 2002).
local function newton_step(lambda, phi, x, y)
  local cp = cos(phi)
  local sp = sin(phi)
  local s2p = sin(2 * phi)
  local sp_sq = sp * sp
  local cp_sq = cp * cp
  local sl = sin(lambda)
  local ch = cos(lambda / 2)
  local sh = sin(lambda / 2)
  local sh_sq = sh * sh
  local gap = 1 - cp_sq * ch * ch          -- 1 - cos^2(angular distance to the centre)
  local arc, inv                           -- distance / sin(distance), 1 / sin^2(distance)
  if gap ~= 0 then
    inv = 1/gap
    arc = acos(cp * ch) * sqrt(inv)
  else
    arc = 0
    inv = 0
  end
  local fx = .5 * (2 * arc * cp * sh + lambda / halfpi) - x
  local fy = .5 * (arc * sp + phi) - y
  local x_l = .5 * inv * (cp_sq * sh_sq + arc * cp * ch * sp_sq) + .5 / halfpi
  local x_p = inv * (sl * s2p / 4 - arc * sp * sh)
  local y_l = .125 * inv * (s2p * sh - arc * sp * cp_sq * sl)
  local y_p = .5 * inv * (sp_sq * ch + arc * sh_sq * cp) + .5
  local det = x_p * y_l - y_p * x_l
  return (fy * x_p - fx * y_p) / det, (fx * y_l - fy * x_l) / det
end

function lens_inverse(x, y)
  if abs(y) >= lens_height/2 then return nil end
  if abs(x) > cut_x and abs(y) > cut_y then return nil end

  local lambda, phi = x, y
  for _ = 1, 25 do
    local dl, dp = newton_step(lambda, phi, x, y)
    lambda = lambda - dl
    phi = phi - dp
    if abs(dl) < eps and abs(dp) < eps then break end
  end

  -- keep only points inside the outline at this latitude
  local outline = lens_forward(latlon_to_ray(phi, pi))
  if abs(x) < abs(outline) then
    return latlon_to_ray(phi, lambda)
  end
  return nil
end
