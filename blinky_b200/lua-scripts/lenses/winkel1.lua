-- Winkel I (forward map only)
max_fov = 360
max_vfov = 180
lens_width = pi * (2/pi + 1)/2 * 2
lens_height = pi
onload = "f_contain"

function lens_forward(x, y, z)
  local lat, lon = ray_to_latlon(x, y, z)
  return lon * (2/pi + cos(lat))/2, lat
end
