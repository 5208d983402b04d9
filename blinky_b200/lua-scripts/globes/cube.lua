-- cube: six 90-degree plates (front, right, left, back, top, bottom).
-- Plate order fixes the rubix tint colour of each face, so it is part of the format.
local function plate(fx, fy, fz, ux, uy, uz)
  return { { fx, fy, fz }, { ux, uy, uz }, 90 }
end

plates = {
  plate( 0,  0,  1,   0, 1,  0),  -- front
  plate( 1,  0,  0,   0, 1,  0),  -- right
  plate(-1,  0,  0,   0, 1,  0),  -- left
  plate( 0,  0, -1,   0, 1,  0),  -- back
  plate( 0,  1,  0,   0, 0, -1),  -- top
  plate( 0, -1,  0,   0, 0,  1),  -- bottom
}
