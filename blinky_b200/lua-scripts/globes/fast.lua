-- fast: two plates looking forward, a sharp 90-degree one inside a coarse 160-degree one.
-- Shows the globe_plate(x,y,z) hook: the script, not the nearest-axis rule, picks the plate.
local SMALL, BIG = 0, 1
local big_fov = 160

plates = {
  { {0,0,1}, {0,1,0}, 90 },
  { {0,0,1}, {0,1,0}, big_fov },
}

function globe_plate(x, y, z)
  if z <= 0 then
    return nil  -- behind the camera: nothing to show
  end
  local dist = 0.5 / tan(big_fov*pi/180/2)
  local size = 2*dist*tan(pi/4)   -- footprint of the small plate on the big one
  local u = x/z*dist
  local v = y/z*dist
  if abs(u) < size/2 and abs(v) < size/2 then
    return SMALL
  end
  return BIG
end
