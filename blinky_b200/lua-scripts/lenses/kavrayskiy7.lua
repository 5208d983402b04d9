-- Kavrayskiy VII (forward map only)
max_fov = 360
max_vfov = 180
lens_width = 3*pi/(2*pi)*sqrt(pi*pi/3)*2
lens_height = pi
onload = "f_contain"

function lens_forward(x, y, z)
  local lat, lon = ray_to_latlon(x, y, z)
  return 3*lon/(2*pi)*sqrt(pi*pi/3 - lat*lat), lat
end
