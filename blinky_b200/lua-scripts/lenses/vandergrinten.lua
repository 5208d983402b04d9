-- Van der Grinten I: the whole sphere inside a circle of radius pi (the old National
-- Geographic world map).  Neither conformal nor equal-area; meridians and parallels
-- are circular arcs.
--
-- forward (Snyder, "Map Projections - A Working Manual", eq. 29-1 .. 29-9):
--   t = asin|2 lat/pi|,  a = |pi/lon - lon/pi| / 2,  g = cos t / (sin t + cos t - 1),
--   p = g (2/sin t - 1),  q = a^2 + g
--   x = +-pi (a (g - p^2) + sqrt(a^2 (g - p^2)^2 - (p^2 + a^2)(g^2 - p^2))) / (p^2 + a^2)
--   y = +-pi (p q - a sqrt((a^2 + 1)(p^2 + a^2) - q^2)) / (p^2 + a^2)
-- inverse: the closed-form cubic solution of the same reference (eq. 29-10 .. 29-21).
onload = "f_contain"
max_vfov = 180
max_fov = 360

local function signed_like(value, reference_product)
  if reference_product < 0 then return -value end
  return value
end

function lens_forward(rx, ry, rz)
  local lat, lon = ray_to_latlon(rx, ry, rz)
  if lat == 0 then return lon, 0 end          -- the equator is a straight line
  local t = asin(abs(2*lat/pi))
  if abs(lat) == pi/2 then                    -- the poles sit on the central meridian
    local polar = pi*tan(t/2)
    return 0, signed_like(polar, polar*lat)
  end
  local a = 0.5*abs(pi/lon - lon/pi)
  local g = cos(t)/(sin(t)+cos(t)-1)
  local p = g*(2/sin(t) - 1)
  local q = a*a+g
  local east = pi*(a*(g-p*p) + sqrt(a*a*(g-p*p)*(g-p*p)-(p*p+a*a)*(g*g-p*p)))/(p*p+a*a)
  local north = pi*(p*q-a*sqrt((a*a+1)*(p*p+a*a) - q*q))/(p*p+a*a)
  return signed_like(east, lon*east), signed_like(north, lat*north)
end

-- constants of the cubic, to 20 digits
local TINY       = 1.e-10
local ONE_THIRD  = .33333333333333333333
local TWO_27THS  = .07407407407407407407
local FOUR_PI_3  = 4.18879020478639098458
local PI_SQ      = 9.86960440108935861869
local TWO_PI_SQ  = 19.73920880217871723738
local HALF_PI_SQ = 4.93480220054467930934

-- the map fills the circle through the image of (lat 0, lon pi)
local rim = lens_forward(latlon_to_ray(0, pi))
lens_width = 2*rim
lens_height = 2*rim

-- longitude from x once r = x^2 + y^2 and the discriminant are known
local function longitude(x, r, discriminant)
  if abs(x) <= TINY then return 0 end
  if discriminant <= 0 then return 0.5 * (r - PI_SQ) / x end
  return 0.5 * (r - PI_SQ + sqrt(discriminant)) / x
end

function lens_inverse(x, y)
  if x*x+y*y > rim*rim then return nil end
  local xx = x*x
  local ay = abs(y)
  if ay < TINY then                           -- on the equator the cubic degenerates
    local disc = xx*xx + TWO_PI_SQ * (xx + HALF_PI_SQ)
    local lon = 0
    if not (abs(x) <= TINY) then lon = 0.5 * (xx - PI_SQ + sqrt(disc)) / x end
    return latlon_to_ray(0, lon)
  end
  local yy = y*y
  local r = xx+yy
  local rr = r*r
  local c1 = -pi*ay*(r+PI_SQ)
  local c3 = rr + (2*pi)*(ay*r+pi*(yy+pi*(ay+pi/2)))
  local c2 = c1 + PI_SQ * (r-3*yy)
  local c0 = pi*ay
  c2 = c2/c3
  local al = c1 / c3 - ONE_THIRD * c2*c2
  local m = 2 *sqrt(-ONE_THIRD*al)
  local d = TWO_27THS*c2*c2*c2+(c0*c0-ONE_THIRD*c2*c1)/c3
  d = 3*d/(al*m)
  local size = abs(d)
  if not (size - TINY <= 1) then return nil end   -- no real root: outside the map
  if size > 1 then                                -- rounding pushed |cos| past 1
    if d > 0 then d = 0 else d = pi end
  else
    d = acos(d)
  end
  local lat = pi * (m*cos(d*ONE_THIRD+FOUR_PI_3) - ONE_THIRD*c2)
  if y < 0 then lat = -lat end
  return latlon_to_ray(lat, longitude(x, r, rr + TWO_PI_SQ * (xx-yy+HALF_PI_SQ)))
end
