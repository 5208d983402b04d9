-- tetra: four plates on the faces of a tetrahedron.
local third = tau/3        -- 120 degrees
local sixth = third / 2    -- 60 degrees

-- tetrahedron measures, with face-centre-to-vertex distance r = 1
local r = 1
local s = 2*r*sin(sixth)          -- edge length
local h = sqrt(s*s-r*r)           -- height over a face
local theta = acos(r/s)
local c = s/2/sin(theta)          -- centre to vertex
local e = r*cos(sixth)            -- face centre to edge
local f = h-c                     -- centre to face

-- field of view that covers a whole face (+1 degree closes the seam at the centre)
local fovr = 2*atan(r/f)
local fovd = fovr * 180 / pi + 1
print(fovd)

local y = e - e*e/(r+e)
local z = -f + h*e/(r+e)

plates = {
  { {0,-y/f,z/f}, {0,-(e-y)/e,(-f-z)/e}, fovd },                                                            -- bottom
  { {y/f*sin(third),-y/f*cos(third),z/f}, {(e-y)/e*sin(third),-(e-y)/e*cos(third),(-f-z)/e}, fovd },        -- right
  { {y/f*sin(-third),-y/f*cos(-third),z/f}, {(e-y)/e*sin(-third),-(e-y)/e*cos(-third),(-f-z)/e}, fovd },    -- left
  { {0,0,-1}, {0,-1,0}, fovd },                                                                             -- back
}
