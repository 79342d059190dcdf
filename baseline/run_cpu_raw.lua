--[[
baseline/run_cpu_raw.lua -- times the UNTOUCHED reference (cpu-raw.lua) with a real LuaJIT, when one exists.

	luajit baseline/run_cpu_raw.lua [size=64] [real=double] [reference_dir=/root/reference]

Prints the reference's own `#iter err` lines followed by one line
	cpu-raw.lua <size> <real> seconds_for_run <seconds> vcycles_per_s <rate>
(run() is two V-cycles, cpu-raw.lua:245). Needs the reference's un-vendored dependencies on package.path
(thenumbernine's lua-ext and lua-image); without them -- or without LuaJIT's ffi -- it prints a line starting with
SKIP and exits 0, so that a harness can always call it. The image this project was built in has no Lua runtime at
all: there the reference's source is executed for CORRECTNESS by oracle/minilua.py (tests/golden/ref_*.npz), and the
CPU timing beside the GPU numbers comes from the C restatement (bench.py cpu_baseline, kind = "port").
For a cross-check of the interpreter: at size 64, real double, the err lines must read 15402.468010923 and
800.681854268.
--]]
local size = tonumber(arg and arg[1]) or 64
local real = (arg and arg[2]) or 'double'
local refdir = (arg and arg[3]) or '/root/reference'

local function skip(why)
	print('SKIP: '..tostring(why))
	os.exit(0)
end

if not pcall(require, 'ffi') then skip('no LuaJIT ffi in this interpreter') end
package.path = refdir..'/?.lua;'..package.path
local ok, MultigridCPURaw = pcall(require, 'cpu-raw')
if not ok then skip('cannot load cpu-raw.lua or one of its dependencies (ext, image): '..tostring(MultigridCPURaw)) end

local multigrid = MultigridCPURaw(size, real)
local t0 = os.clock()
multigrid:run()
local dt = os.clock() - t0
print(('cpu-raw.lua %d %s seconds_for_run %.6f vcycles_per_s %.6f'):format(size, real, dt, 2 / dt))
