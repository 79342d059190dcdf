#!/usr/bin/env luajit
--[[
Timing driver for the `cuda` column, writing the artefact of the reference's test/test.lua in its format:
a tab-separated table `cpu-vs-gpu.txt` whose first line is `#size<TAB>col1<TAB>col2...` and whose rows are
`size<TAB>best time of col1<TAB>...` (test/test.lua:16-33,45-65; time = os.clock() around multigrid:run()).

  luajit test-cuda.lua [log2 of the first size] [log2 of the last size] [cpudepth] [col,col,...]

Columns default to {'cuda'}; any module the reference ships can be named next to it (e.g. cuda,cpu-raw) when its
dependencies are installed. To add the column to the reference's own driver instead, put 'cuda' into the `cols`
table at test/test.lua:8-14: the class has the constructor and run() that driver calls (INTEGRATION.md).
Plotting (gnuplot, test/test.lua:67-76) is left to the reference: this file only writes the table.
--]]
local bit = require 'bit'

local firstLog2 = tonumber(arg and arg[1]) or 5		-- test/test.lua:45 runs 2^5 only
local lastLog2 = tonumber(arg and arg[2]) or firstLog2
local cpudepth = tonumber(arg and arg[3]) or 3			-- test/test.lua:42
local cols = {}
for name in ((arg and arg[4]) or 'cuda'):gmatch('[^,]+') do cols[#cols+1] = name end
local tries = tonumber(os.getenv'MGPOISSON_TRIES') or 1	-- test/test.lua:44
local outName = os.getenv'MGPOISSON_TSV' or 'cpu-vs-gpu.txt'

local out = assert(io.open(outName, 'w'))
local function emit(text)		-- to the terminal and to the file, like the reference's write()
	io.write(text)
	out:write(text)
	out:flush()
end

local header = {'#size'}
for _,col in ipairs(cols) do header[#header+1] = col end
emit(table.concat(header, '\t')..'\n')

for log2size = firstLog2, lastLog2 do
	local size = bit.lshift(1, log2size)
	local row = {tostring(size)}
	for _,col in ipairs(cols) do
		local class = require('multigrid-poisson.'..col)
		local best = math.huge
		for _ = 1, tries do
			local solver = class(size, nil, cpudepth)
			local t0 = os.clock()
			solver:run()
			best = math.min(best, os.clock() - t0)
		end
		row[#row+1] = tostring(best)
	end
	emit(table.concat(row, '\t')..'\n')
end
out:close()
