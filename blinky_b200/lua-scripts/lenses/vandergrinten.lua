-- Van der Grinten I: the sphere in a circle.
max_fov = 360
max_vfov = 180
onload = "f_contain"

function lens_forward(x, y, z)
  local lat, lon = ray_to_latlon(x, y, z)
  if lat == 0 then
    return lon, 0
  end
  local t = asin(abs(2*lat/pi))
  if abs(lat) == pi/2 then
    local y2 = pi*tan(t/2)
    if y2*lat < 0 then
      y2 = -y2
    end
    return 0,y2
  end
  local a = 0.5*abs(pi/lon - lon/pi)
  local g = cos(t)/(sin(t)+cos(t)-1)
  local p = g*(2/sin(t) - 1)
  local q = a*a+g

  local px = pi*(a*(g-p*p) + sqrt(a*a*(g-p*p)*(g-p*p)-(p*p+a*a)*(g*g-p*p)))/(p*p+a*a)
  local py = pi*(p*q-a*sqrt((a*a+1)*(p*p+a*a) - q*q))/(p*p+a*a)

  if lon*px < 0 then
    px = -px
  end
  if lat*py < 0 then
    py = -py
  end
  return px, py
end

local TOL      = 1.e-10
local THIRD    = .33333333333333333333
local C2_27    = .07407407407407407407
local PI4_3    = 4.18879020478639098458
local PISQ     = 9.86960440108935861869
local TPISQ    = 19.73920880217871723738
local HPISQ    = 4.93480220054467930934

local rim = lens_forward(latlon_to_ray(0, pi))
lens_height = 2*rim
lens_width = 2*rim

function lens_inverse(x, y)
  if x*x+y*y > rim*rim then
    return nil
  end
  local lat, lon
  local t, c0, c1, c2, c3, al, r2, r, m, d, ay, x2, y2

  x2 = x*x
  ay = abs(y)
  if ay < TOL then
    -- on the equator
    lat = 0
    t = x2*x2 + TPISQ * (x2 + HPISQ)
    if abs(x) <= TOL then
      lon = 0
    else
      lon = 0.5 * (x2 - PISQ + sqrt(t)) / x
    end
    return latlon_to_ray(lat, lon)
  end

  y2 = y*y
  r = x2+y2
  r2 = r*r
  c1 = -pi*ay*(r+PISQ)
  c3 = r2 + (2*pi)*(ay*r+pi*(y2+pi*(ay+pi/2)))
  c2 = c1 + PISQ * (r-3*y2)
  c0 = pi*ay
  c2 = c2/c3
  al = c1 / c3 - THIRD * c2*c2
  m = 2 *sqrt(-THIRD*al)
  d = C2_27*c2*c2*c2+(c0*c0-THIRD*c2*c1)/c3
  d = 3*d/(al*m)
  t = abs(d)
  if not (t - TOL <= 1) then
    return nil
  end
  if t > 1 then
    if d > 0 then
      d = 0
    else
      d = pi
    end
  else
    d = acos(d)
  end
  lat = pi * (m*cos(d*THIRD+PI4_3) - THIRD*c2)
  if y < 0 then
    lat = -lat
  end
  t = r2 + TPISQ * (x2-y2+HPISQ)
  if abs(x) <= TOL then
    lon = 0
  elseif t <= 0 then
    lon = 0.5 * (r - PISQ) / x
  else
    lon = 0.5 * (r - PISQ + sqrt(t)) / x
  end
  return latlon_to_ray(lat, lon)
end
