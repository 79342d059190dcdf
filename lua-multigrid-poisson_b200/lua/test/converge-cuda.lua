#!/usr/bin/env luajit
--[[
The experiment of the reference's test/converge-multigrid-vs-krylov.lua on the CUDA library: for each grid size the
multigrid solver in its cpu.lua form (table constructor, errorCallback per cycle, :solve(); coarse corrections start
from zero every cycle, cpu.lua:138) records |psi|_inf per cycle (:20-30), then conjugate gradient on the same operator
with x0 = -f, b = f records |x|_inf per iteration (:38-69); the smallest recorded value is subtracted from everything
(:71-85) and the table is written to converge/<size>.txt, one row per iteration, columns multigrid<TAB>conjgrad,
`nan` where a column has no entry (:79-89).

  luajit converge-cuda.lua [maxiter] [size size ...]      (default: 1000 cycles, sizes 4 8 16 32 64 128)

Differences from the reference, both forced: the Krylov solver is the library's mg_cg (the reference's `solver.conjgrad`
is an un-vendored dependency; same recurrence, same stop test err < epsilon), and there is no plot.
--]]
local MultigridCUDA = require 'multigrid-poisson.cuda'

local maxiter = tonumber(arg and arg[1]) or 1000		-- cpu.lua:22
local sizes = {}
for i = 2, (arg and #arg or 0) do sizes[#sizes+1] = tonumber(arg[i]) end
if #sizes == 0 then sizes = {4, 8, 16, 32, 64, 128} end		-- converge-multigrid-vs-krylov.lua:15
local epsilon = 1e-20										-- :12
local outDir = os.getenv'MGPOISSON_CONVERGE_DIR' or 'converge'
os.execute('mkdir -p "'..outDir..'"')

local function isfinite(x) return x == x and x ~= math.huge and x ~= -math.huge end

for _,size in ipairs(sizes) do
	print('solving for size '..size)
	local rows = {}			-- rows[iter] = {multigrid value, conjgrad value}
	local mg
	mg = MultigridCUDA{
		size = size,
		maxiter = maxiter,
		epsilon = epsilon,
		errorCallback = function(iter, err)
			rows[iter] = {mg.psi:normLInf()}
		end,
	}
	mg:solve()

	-- conjugate gradient from the experiment's starting point x0 = -f (what initCells leaves in psi)
	mg:initCells()
	local n, errs, linf = mg:conjgrad(size * size * 4, epsilon)
	for iter = 1, n do
		rows[iter] = rows[iter] or {}
		rows[iter][2] = linf[iter-1]
	end

	local count = 0
	for iter in pairs(rows) do count = math.max(count, iter) end
	local lowest = math.huge
	for iter = 1, count do
		rows[iter] = rows[iter] or {}
		for col = 1, 2 do
			local v = rows[iter][col]
			if v == nil then rows[iter][col] = 0/0 elseif isfinite(v) then lowest = math.min(lowest, v) end
		end
	end
	local lines = {}
	for iter = 1, count do
		local cells = {}
		for col = 1, 2 do cells[col] = tostring(rows[iter][col] - lowest) end
		lines[iter] = table.concat(cells, '\t')
	end
	local out = assert(io.open(outDir..'/'..size..'.txt', 'w'))
	out:write(table.concat(lines, '\n'))
	out:close()
end
