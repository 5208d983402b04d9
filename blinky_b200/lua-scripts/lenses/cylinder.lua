-- Central cylindrical: the globe projected from its centre onto a wrapped cylinder,
-- x = longitude, y = tan(latitude).  Endless vertically, so only a width is given
-- and the default zoom covers the screen.
onload = "f_cover"
lens_width = 2*pi
max_vfov = 180
max_fov = 360

function lens_inverse(x, y)
  if abs(x) > pi then return nil end   -- one turn around the cylinder
  return latlon_to_ray(atan(y), x)
end

function lens_forward(rx, ry, rz)
  local lat, lon = ray_to_latlon(rx, ry, rz)
  return lon, tan(lat)
end
