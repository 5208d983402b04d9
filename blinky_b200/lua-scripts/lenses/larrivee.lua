-- Larrivee (1988), a compromise world map with bulging polar regions.  Forward map only.
--
--   x = lon (1 + sqrt(cos lat)) / 2
--   y = lat / (cos(lat/2) cos(lon/6))
onload = "f_contain"
max_vfov = 180
max_fov = 360
lens_height = pi/2 / cos(pi/2/2) * 2   -- y at the pole on the central meridian, doubled
lens_width = 2*pi

local function project(lat, lon)
  return (0.5 + 0.5*sqrt(cos(lat)))*lon, lat / (cos(lat/2)*cos(lon/6))
end

function lens_forward(rx, ry, rz)
  return project(ray_to_latlon(rx, ry, rz))
end
