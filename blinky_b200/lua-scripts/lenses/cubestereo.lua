-- Stereographic view of the cube map (rays pushed onto the cube first).
max_fov = 270
max_vfov = 270
onload = "f_fov 180"

local function onto_cube(x, y, z)
  local mx, my, mz = abs(x), abs(y), abs(z)
  local m = mz
  if mx >= my and mx >= mz then
    m = mx
  elseif my >= mx and my >= mz then
    m = my
  end
  return x / m, y / m, z / m
end

function lens_forward(rx, ry, rz)
  local x, y, z = onto_cube(rx, ry, rz)
  return x/(z+1)*2, y/(z+1)*2
end

function lens_inverse(x, y)
  local rx, ry, rz
  local mx = abs(x)
  local my = abs(y)
  local z = 2
  if mx <= 1 and my <= 1 then
    rx = x
    ry = y
    rz = z-1
  elseif mx > my then
    rx = x / mx
    ry = y / mx
    rz = z / mx-1
  else
    rx = x / my
    ry = y / my
    rz = z / my-1
  end
  local len = sqrt(rx*rx+ry*ry+rz*rz)
  return rx/len, ry/len, rz/len
end
