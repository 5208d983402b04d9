-- Winkel II: the mean of an elliptical (Apian II style) and an equirectangular map.
--   x = lon/2 (2/pi + sqrt(pi^2 - 4 lat^2)/pi),   y = lat.        Forward map only.
onload = "f_contain"
lens_height = pi
lens_width = pi/2*(2/pi+1)*2
max_vfov = 180
max_fov = 360

local function project(lat, lon)
  local ellipse = sqrt(pi*pi - 4*lat*lat)/pi   -- half-width of the elliptical component, 1 at the equator
  return lon/2*(2/pi + ellipse), lat
end

function lens_forward(rx, ry, rz)
  return project(ray_to_latlon(rx, ry, rz))
end
