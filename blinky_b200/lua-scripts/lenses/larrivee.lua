-- Larrivee projection (forward map only).
max_fov = 360
max_vfov = 180
lens_width = 2*pi
lens_height = pi/2 / cos(pi/2/2) * 2
onload = "f_contain"

function lens_forward(x, y, z)
  local lat, lon = ray_to_latlon(x, y, z)
  local px = (0.5 + 0.5*sqrt(cos(lat)))*lon
  local py = lat / (cos(lat/2)*cos(lon/6))
  return px, py
end
