-- cube_edge: the cube turned 45 degrees so that an edge faces the viewer.
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

-- rotate a vector in place: first about the vertical axis, nothing else
local function turn(v)
  local x, z = v[1], v[3]
  v[1] = x*cos(a)-z*sin(a)
  v[3] = x*sin(a)+z*cos(a)

end

for i = 1, 6 do
  turn(plates[i][1])  -- forward
  turn(plates[i][2])  -- up
end
