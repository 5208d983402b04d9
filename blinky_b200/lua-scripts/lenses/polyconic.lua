-- American polyconic: every parallel is the arc it would be on its own tangent cone.
-- Forward map only; no natural frame, so it opens at a 360 degree fit.
--
--   E = lon sin lat,   x = cot lat sin E,   y = lat + cot lat (1 - cos E)
--   (on the equator the cones degenerate: x = lon, y = 0)
onload = "f_fov 360"
max_vfov = 180
max_fov = 360

local function project(lat, lon)
  if lat == 0 then return lon, 0 end
  return 1/tan(lat)*sin(lon*sin(lat)), lat + 1/tan(lat)*(1 - cos(lon*sin(lat)))
end

function lens_forward(rx, ry, rz)
  return project(ray_to_latlon(rx, ry, rz))
end
