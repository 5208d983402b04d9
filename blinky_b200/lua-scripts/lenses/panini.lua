-- Pannini ("Panini") projection, compression parameter d = 1: a cylindrical
-- stereographic view.  Verticals stay straight and radial lines through the centre
-- stay straight, so very wide interior views (up to 360 degrees across) still look
-- natural.  See Sharpless, Postle & German, "Pannini: A New Projection for Rendering
-- Wide Angle Perspective Images" (2010).
--
--   forward: S = (d+1)/(d + cos lon),  x = S sin lon,  y = S tan lat
--   inverse: k = x^2/(d+1)^2,  cos lon = (-k d + sqrt(k^2 d^2 - (k+1)(k d^2 - 1)))/(k+1)
onload = "f_fov 180"
max_vfov = 180
max_fov = 360

local d = 1

-- cosine of the longitude of the image column x
local function cos_lon(x)
  local k = x*x/((d+1)*(d+1))
  local dscr = k*k*d*d - (k+1)*(k*d*d-1)
  return (-k*d+sqrt(dscr))/(k+1)
end

function lens_inverse(x, y)
  local clon = cos_lon(x)
  local S = (d+1)/(d+clon)
  local lon = atan2(x,S*clon)
  return latlon_to_ray(atan2(y,S), lon)
end

function lens_forward(rx, ry, rz)
  local lat, lon = ray_to_latlon(rx, ry, rz)
  local S = (d+1)/(d+cos(lon))
  return S*sin(lon), S*tan(lat)
end
