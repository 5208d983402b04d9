-- Eckert I: pseudocylindrical, straight meridians broken at the equator; poles are
-- lines half as long as the equator.  Only the forward map is defined, so the
-- lensmap is built by projecting the globe's texels onto the screen.
--
--   x = FC lon (1 - |lat|/pi),   y = FC lat,   FC = 2 sqrt(2 / (3 pi))
onload = "f_contain"
max_vfov = 180
max_fov = 360

local FC = 0.92131773192356127802   -- 2 sqrt(2/(3 pi))
local RP = 0.31830988618379067154   -- 1/pi
lens_height = FC * pi
lens_width = FC * pi * 2

local function project(lat, lon)
  return FC * lon * (1 - RP * abs(lat)), FC * lat
end

function lens_forward(rx, ry, rz)
  return project(ray_to_latlon(rx, ry, rz))
end
