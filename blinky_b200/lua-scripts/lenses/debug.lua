-- debug: shows every plate of the current globe side by side (uses numplates and plate_to_ray).
local rows, cols
if numplates == 4 then
  rows, cols = 2, {2,2}
elseif numplates == 5 then
  rows, cols = 2, {3,2}
elseif numplates == 6 then
  rows, cols = 2, {3,3}
else
  rows, cols = 1, {numplates}
end
local widest = math.max(table.unpack(cols))

lens_width = widest
lens_height = rows
onload = "f_contain"

-- (cell index, position in cell) or nil,nil when n lies outside [0,count)
local function cell(n, count)
  local i, f = math.modf(n)
  if n < 0 or n >= count then
    return nil, nil
  end
  return i, f
end

function lens_inverse(x, y)
  local r, v = cell(-y+rows/2, rows)
  if r == nil then
    return nil
  end
  local c, u = cell(x+cols[r+1]/2, cols[r+1])
  if c == nil then
    return nil
  end
  local plate = c
  local i = 0
  while i < r do
    plate = plate + cols[i+1]
    i = i + 1
  end
  return plate_to_ray(plate, u, v)
end
