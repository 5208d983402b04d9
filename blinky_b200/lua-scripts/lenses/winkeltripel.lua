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

function lens_inverse(x, y)
  if abs(y) >= lens_height/2 then
    return nil
  end
  if abs(x) > cut_x and abs(y) > cut_y then
    return nil
  end

  local lambda = x
  local phi = y
  for iter = 1, 25 do
    local cosphi = cos(phi)
    local sinphi = sin(phi)
    local sin_2phi = sin(2 * phi)
    local sin2phi = sinphi * sinphi
    local cos2phi = cosphi * cosphi
    local sinlambda = sin(lambda)
    local coslambda_2 = cos(lambda / 2)
    local sinlambda_2 = sin(lambda / 2)
    local sin2lambda_2 = sinlambda_2 * sinlambda_2
    local C = 1 - cos2phi * coslambda_2 * coslambda_2
    local E, F
    if C ~= 0 then
      F = 1/C
      E = acos(cosphi * coslambda_2) * sqrt(F)
    else
      E = 0
      F = 0
    end
    local fx = .5 * (2 * E * cosphi * sinlambda_2 + lambda / halfpi) - x
    local fy = .5 * (E * sinphi + phi) - y
    local dxdl = .5 * F * (cos2phi * sin2lambda_2 + E * cosphi * coslambda_2 * sin2phi) + .5 / halfpi
    local dxdp = F * (sinlambda * sin_2phi / 4 - E * sinphi * sinlambda_2)
    local dydl = .125 * F * (sin_2phi * sinlambda_2 - E * sinphi * cos2phi * sinlambda)
    local dydp = .5 * F * (sin2phi * coslambda_2 + E * sin2lambda_2 * cosphi) + .5
    local den = dxdp * dydl - dydp * dxdl
    local dl = (fy * dxdp - fx * dydp) / den
    local dp = (fx * dydl - fy * dxdl) / den
    lambda = lambda - dl
    phi = phi - dp
    if abs(dl) < eps and abs(dp) < eps then
      break
    end
  end

  -- keep only points inside the outline at this latitude
  local x0 = lens_forward(latlon_to_ray(phi, pi))
  if abs(x) < abs(x0) then
    return latlon_to_ray(phi, lambda)
  end
  return nil
end
