-- cube_corner: the cube turned so that a corner faces the viewer.
local function plate(fx, fy, fz, ux, uy, uz)
  return { { fx, fy, fz }, { ux, uy, uz }, 90 }
end

plates = {
  plate( 0,  0,  1,   0, 1,  0),
  plate( 1,  0,  0,   0, 1,  0),
  plate(-1,  0,  0,   0, 1,  0),
  plate( 0,  0, -1,   0, 1,  0),
  plate( 0,  1,  0,   0, 0, -1),
  plate( 0, -1,  0,   0, 0,  1),
}

local a = pi/4

-- rotate a vector in place: first about the vertical axis, then about the horizontal one
local function turn(v)
  local x, z = v[1], v[3]
  v[1] = x*cos(a)-z*sin(a)
  v[3] = x*sin(a)+z*cos(a)
  local y
  y, z = v[2], v[3]
  v[2] = y*cos(a)-z*sin(a)
  v[3] = y*sin(a)+z*cos(a)
end

for i = 1, 6 do
  turn(plates[i][1])  -- forward
  turn(plates[i][2])  -- up
end
