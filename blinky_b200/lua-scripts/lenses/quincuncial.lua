-- Peirce quincuncial: the sphere on a square, via Jacobi elliptic functions.
local eps = 0.0001
local halfpi = pi/2

local function safe_sqrt(x)
  if x > 0 then
    return sqrt(x)
  end
  return 0
end

-- Jacobi elliptic functions sn, cn, dn (and the amplitude) of u with parameter m,
-- by the arithmetic-geometric-mean descent (same scheme as Matlab's ellipj).
local function ellipj(u, m)
  local ai, b, phi, t, twon
  if m < eps then
    t = sin(u)
    b = cos(u)
    ai = .25 * m * (u - t * b)
    return t - ai * b, b + ai * t, 1 - .5 * m * t * t, u - ai
  end
  if m >= 1 - eps then
    ai = .25 * (1 - m)
    b = cosh(u)
    t = tanh(u)
    phi = 1 / b
    twon = b * sinh(u)
    return t + ai * (twon - u) / (b * b),
           phi - ai * t * phi * (twon - u),
           phi + ai * t * phi * (twon + u),
           2 * atan(exp(u)) - halfpi + ai * (twon - u) / b
  end

  local a = {1, 0, 0, 0, 0, 0, 0, 0, 0}
  local c = {sqrt(m), 0, 0, 0, 0, 0, 0, 0, 0}
  local i = 1
  b = sqrt(1 - m)
  twon = 1
  while abs(c[i] / a[i]) > eps and i < 9 do
    ai = a[i]
    i = i+1
    c[i] = .5 * (ai - b)
    a[i] = .5 * (ai + b)
    b = safe_sqrt(ai * b)
    twon = twon*2
  end

  phi = twon * a[i] * u
  repeat
    b = phi
    t = c[i] * sin(b) / a[i]
    phi = .5 * (asin(t) + phi)
    i = i-1
  until i == 1

  t = cos(phi)
  return sin(phi), t, t / cos(phi - b), phi
end

local sqrt2 = sqrt(2)
local sqrt22 = sqrt2/2
local m = 1/2
local ke = 1.85407467730137  -- complete elliptic integral K(1/2)

-- point of the unit square (corners at +-1) -> latitude, longitude
-- (Fong & Vogel, "Warping Peirce Quincuncial Panoramas", appendix A)
local function square_to_latlon(x, y)
  local xpr = ke*(sqrt22*x-sqrt22*y)/sqrt2+ke
  local ypr = ke*(sqrt22*x+sqrt22*y)/sqrt2
  local x1, y1
  if abs(ypr) < eps then
    local _, cn = ellipj(xpr, m)
    x1 = cn
    y1 = 0.0
  else
    local s, c, d = ellipj(xpr, m)
    local s1, c1, d1 = ellipj(ypr, 1-m)
    local delta = c1^2 + m*s^2*s1^2
    x1 = (c*c1)/delta
    y1 = -(s*d*s1*d1)/delta
  end
  local lon = atan2(y1,x1)
  local lat = 2*atan2(sqrt(x1*x1+y1*y1),1)-halfpi
  return lat, lon
end

lens_height = 2*sqrt2
lens_width = 2*sqrt2
onload = "f_contain"

local function rotate(a, b, angle)
  local c = cos(angle)
  local s = sin(angle)
  return a*c - b*s, a*s + b*c
end

-- The square is first unfolded into a 4x2 strip: front hemisphere on the left
-- half, back hemisphere (its four corner triangles reassembled) on the right.
local function strip_to_ray(x, y)
  if abs(x) > 2 or abs(y) > 1 then
    return nil
  end
  x = x+1
  local lat, lon = square_to_latlon(x, y)
  local x0, y0, z0 = latlon_to_ray(lat, -lon)
  return x0, z0, -y0   -- swing the south pole round to the view centre
end

function lens_inverse(x, y)
  if abs(x) > sqrt2 or abs(y) > sqrt2 then
    return nil
  end
  local x0, y0
  if abs(x)+abs(y) < sqrt2 then      -- inner diamond: front hemisphere
    x0, y0 = rotate(x,y,pi/4)
    x0 = x0-1
  elseif x>0 and y<0 then            -- lower right corner
    x0, y0 = rotate(x,y,pi/4)
    x0 = x0-1
  elseif x<0 and y>0 then            -- upper left corner
    x0, y0 = rotate(x,y,pi/4)
    x0 = x0+3
  elseif x<0 and y<0 then            -- lower left corner
    x0, y0 = rotate(x,y,pi/4+pi)
    x0, y0 = x0+1, y0-2
  else                               -- upper right corner
    x0, y0 = rotate(x,y,pi/4+pi)
    x0, y0 = x0+1, y0+2
  end
  return strip_to_ray(x0, y0)
end
