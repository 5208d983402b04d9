-- American polyconic projection (forward map only).
max_fov = 360
max_vfov = 180
onload = "f_fov 360"

function lens_forward(x, y, z)
  local lat, lon = ray_to_latlon(x, y, z)
  if lat == 0 then
    return lon, 0
  end
  local px = 1/tan(lat)*sin(lon*sin(lat))
  local py = lat + 1/tan(lat)*(1 - cos(lon*sin(lat)))
  return px, py
end
