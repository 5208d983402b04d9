-- Wagner VI: pseudocylindrical compromise, meridians are arcs of ellipses.
--   x = lon sqrt(1 - 3 (lat/pi)^2),   y = lat.            Forward map only.
onload = "f_contain"
lens_height = pi
lens_width = pi*2
max_vfov = 180
max_fov = 360

local function project(lat, lon) return lon*sqrt(1-3*lat*lat/(pi*pi)), lat end

function lens_forward(rx, ry, rz)
  return project(ray_to_latlon(rx, ry, rz))
end
