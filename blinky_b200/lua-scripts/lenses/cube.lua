-- Unfolded cube: the six faces laid out as a 4x3 cross.
local cols = 4
local rows = 3

lens_width = cols
lens_height = rows
max_fov = 360
max_vfov = 180
onload = "f_contain"

-- split a coordinate into (cell index, position inside the cell)
local function cell(n)
  local i, f = math.modf(n)
  if n < 0 then
    return i-1, f+1
  end
  return i, f
end

function lens_inverse(x, y)
  x = x - 0.5
  local r, v = cell(-y+rows/2)
  local c, u = cell(x+cols/2)
  u = u - 0.5
  v = v - 0.5
  v = -v

  if r < 0 or r >= rows or c < -1 or c >= cols then
    return nil
  end
  if (r == 0 or r == 2) and not (c == 1) then
    return nil  -- only the middle column has a top and a bottom
  end

  if r == 0 then return u,0.5,-v end        -- top
  if r == 2 then return u,-0.5,v end        -- bottom
  if c == 0 then return -0.5,v,u end        -- left
  if c == 1 then return u,v,0.5 end         -- front
  if c == 2 then return 0.5,v,-u end        -- right
  if c == 3 or c == -1 then return -u,v,-0.5 end  -- back (wraps around)
  return nil
end

-- only good enough to place the FOV marks
function lens_forward(x, y, z)
  local ax, ay, az = abs(x), abs(y), abs(z)
  local m = math.max(ax,ay,az)
  local u, v
  if m == ax then
    if x > 0 then
      u = -z/x*0.5
      v = y/x*0.5
      return 1+u,v
    else
      u = z/-x*0.5
      v = y/-x*0.5
      return -1+u,v
    end
  elseif m == ay then
    if y > 0 then
      u = x/y*0.5
      v = -z/y*0.5
      return u,1+v
    else
      u = x/-y*0.5
      v = z/-y*0.5
      return u,-1+v
    end
  elseif m == az then
    if z > 0 then
      u = x/z*0.5
      v = y/z*0.5
      return u,v
    else
      u = -x/-z*0.5
      v = y/-z*0.5
      if u > 0 then
        return -2+u,v
      else
        return 2+u,v
      end
    end
  end
end
