-- trism: triangular prism, three 120-degree side plates plus top and bottom caps.
local c, s = cos(pi/6), sin(pi/6)

plates = {
  { { -c, 0, s }, { 0, 1, 0 }, 120 },   -- left
  { {  c, 0, s }, { 0, 1, 0 }, 120 },   -- right
  { {  0, 0, -1 }, { 0, 1, 0 }, 120 },  -- back
  { {  0, 1, 0 }, { 0, 0, -1 }, 128 },  -- top
  { {  0, -1, 0 }, { 0, 0, -1 }, 128 }, -- bottom
}
