-- Kavrayskiy VII: low-distortion compromise world map (the Soviet atlas standard).
--   x = 3 lon / (2 pi) * sqrt(pi^2/3 - lat^2),   y = lat.          Forward map only.
onload = "f_contain"
lens_height = pi
lens_width = 3*pi/(2*pi)*sqrt(pi*pi/3)*2
max_vfov = 180
max_fov = 360

local function project(lat, lon) return 3*lon/(2*pi)*sqrt(pi*pi/3 - lat*lat), lat end

function lens_forward(rx, ry, rz)
  return project(ray_to_latlon(rx, ry, rz))
end
