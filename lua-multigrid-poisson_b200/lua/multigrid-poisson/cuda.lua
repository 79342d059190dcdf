--[[
multigrid-poisson/cuda.lua -- LuaJIT class that puts libmgpoisson.so (B200, sm_100a CUDA)
behind the solver interface of thenumbernine/lua-multigrid-poisson:

	local cl = require 'multigrid-poisson.cuda'      -- test/test.lua:53
	local multigrid = cl(size, real, cpuDepth)        -- test/test.lua:54
	multigrid:run()                                   -- test/test.lua:56

so adding 'cuda' to `cols` in test/test.lua:8-14 is the whole integration. It mirrors
MultigridCPURaw (cpu-raw.lua:118-258) / MultigridGPU (gpu.lua:18-375): fields size, real,
smooth, accuracy, debugging; methods init, run, twoGrid, inPlaceIterativeSolver; buffers
f, psi, psiOld, errorBuf, tmpU, rs[L], Rs[L], vs[L], Vs[L] (device pointers).

No LuaJIT exists in the build environment; this file is executed from source by the project's
own Lua interpreter with an ffi shim (tests/test_lua_binding.py; both are test infrastructure
outside this package), and the same C ABI is exercised by the Python mirror
lua-multigrid-poisson_b200/__init__.py. Kept declarative on purpose.
The ffi.cdef text below is the MGPOISSON_CDEF block of include/mgpoisson.h, verbatim.
--]]
local ffi = require 'ffi'
local class = require 'ext.class'
local math = require 'ext.math'

ffi.cdef[[
typedef struct mg_ctx mg_ctx;

/* real_kind: storage type / arithmetic type of every field.
 *  MG_REAL_F64       double / double   cpu-raw.lua default (cpu-raw.lua:143), gpu.lua with fp64
 *  MG_REAL_F32       float  / float    gpu.lua on an fp32-only device (gpu.lua:32,39)
 *  MG_REAL_F32_ACC64 float  / double   cpu-raw.lua with real='float' (LuaJIT evaluates in double)
 */
enum { MG_REAL_F64 = 0, MG_REAL_F32 = 1, MG_REAL_F32_ACC64 = 2 };

/* which: the reference's public buffers. `level` is the grid width L of that level
 * (ignored for the five full-size buffers). */
enum {
    MG_BUF_F = 0, MG_BUF_PSI = 1, MG_BUF_PSIOLD = 2, MG_BUF_ERRORBUF = 3, MG_BUF_TMPU = 4,
    MG_BUF_r = 5, MG_BUF_R = 6, MG_BUF_v = 7, MG_BUF_V = 8
};

/* mode of mg_vcycle / mg_step / mg_run:
 *  MG_MODE_FUSED   production path: temporally blocked smoother, residual+restriction and
 *                  prolongation+add fused into the smoother passes, persistent kernel for
 *                  the small levels, ping-pong instead of copy-back. rs[L]/vs[L] are never
 *                  materialised. Per-point arithmetic is identical to MG_MODE_REFSEQ.
 *  MG_MODE_REFSEQ  one kernel per reference operator in the reference's order, including the
 *                  tmpU copy-back (cpu-raw.lua:176-237); materialises rs[L], vs[L]; supports
 *                  the stage trace. This is the `debugging = true` path. */
enum { MG_MODE_FUSED = 0, MG_MODE_REFSEQ = 1 };

enum {
    MG_OK = 0, MG_EINVAL = -1, MG_ECUDA = -2, MG_ENOMEM = -3, MG_ESTATE = -4, MG_EUNSUPPORTED = -5
};

/* ---- lifetime (replaces Class(size, real, cpuDepth): cpu-raw.lua:142-174, gpu.lua:26-245).
 * Allocates the grid hierarchy in one device arena, zero-fills it once, and runs initCells
 * (point source, psi = -f). dim = 2 is the reference; dim = 3 is this project's extension.
 * smooth <= 0 selects the reference default 7 (cpu-raw.lua:123). device < 0 = current device. */
int mg_create(int dim, int size, int real_kind, int smooth, int device, mg_ctx **out);
int mg_destroy(mg_ctx *ctx);
const char *mg_last_error(mg_ctx *ctx);      /* ctx may be NULL: error of a failed mg_create */
const char *mg_version(void);

/* ---- knobs */
int mg_set_mode(mg_ctx *ctx, int mode);
int mg_set_stream(mg_ctx *ctx, void *cuda_stream);   /* borrow a cudaStream_t (NULL = own) */
/* fused-path tuning: tb = Jacobi sweeps per smoother pass of the streaming (TMA) smoother
 * (1..4; 0 = untiled one-sweep kernels), small_L = largest level width handled by the
 * persistent small-level kernel, use_graph = replay the V-cycle from a CUDA graph.
 * Negative = keep. None of these changes a single bit of any result. */
int mg_set_tuning(mg_ctx *ctx, int tb, int small_L, int use_graph);
/* named integer options: "tb", "small_L", "graph", "stream_min_L" (smallest level width
 * given to the streaming smoother), "tz" (planes per CTA of the streaming smoother; 0 = auto), "tb2" (2-D: sweeps per pass of the
 * warp-streaming smoother, 0..7), "warp2d_min_L", "ty" (2-D: rows per warp work item), "slab_p2p" (see below), "tma_promo" */
int mg_set_option(mg_ctx *ctx, const char *name, int value);
/* Relaxation weight of the Jacobi smoother, u + omega (J(u) - u). The reference is omega = 1 (cpu-raw.lua:34-44,
 * 176-184: dest = (f - askew) / adiag, no weight) and that is the default, bit-identical to the reference path. Any other
 * value is an EXTENSION (0 < omega < 2; e.g. 6/7 in 3-D, 4/5 in 2-D damp the mode the reference leaves undamped),
 * offered so that a converging, labelled run can be timed next to the reference's; single GPU only. */
int mg_set_omega(mg_ctx *ctx, double omega);
int mg_get_info(mg_ctx *ctx, int *dim, int *size, int *real_kind, int *smooth, int *nlevels,
                uint64_t *arena_bytes);

/* ---- data movement (replaces enqueueReadBuffer / enqueueWriteBuffer and `.buffer`) */
int mg_init_cells(mg_ctx *ctx);                               /* cpu-raw.lua:8-20,173 */
int mg_zero_corrections(mg_ctx *ctx);                         /* Vs[*] = 0 (cpu.lua:138 variant) */
int mg_upload(mg_ctx *ctx, int which, int level, const void *host, size_t bytes);
int mg_download(mg_ctx *ctx, int which, int level, void *host, size_t bytes);
void *mg_device_ptr(mg_ctx *ctx, int which, int level);       /* NULL if not materialised */
void *mg_host_alloc(size_t bytes);                            /* pinned host memory for upload/download */
int mg_host_free(void *p);

/* ---- the hot path */
int mg_vcycle(mg_ctx *ctx);                                   /* twoGrid(1/size, psi, f, size) */
int mg_vcycle_async(mg_ctx *ctx);                             /* same, no stream sync on return */
int mg_synchronize(mg_ctx *ctx);
int mg_step(mg_ctx *ctx, double *err);                        /* cpu-raw.lua:246-254 */
int mg_run(mg_ctx *ctx, int max_cycles, double accuracy, double *errs, int *n_done);
                                                              /* cpu-raw.lua:239-258 */
/* host-buffer entry point: upload f and psi, one mg_step, download psi. */
int mg_step_host(mg_ctx *ctx, const void *f_host, void *psi_host, double *err);
/* n independent problems in host memory (pinned: mg_host_alloc), each through mg_step_host's sequence, pipelined: while
 * problem i's cycle runs, one copy engine uploads problem i+1 and the other downloads problem i-1 (cpu-gpu.lua:26-48 does
 * these transfers one after the other around its device work). Same results as n calls of mg_step_host, bit for bit;
 * psi_hosts[i] must be distinct buffers (f_hosts[i] may repeat); errs[i] = that problem's err (may be NULL). */
int mg_step_host_batch(mg_ctx *ctx, int n, const void **f_hosts, void **psi_hosts, double *errs);
/* true residual RMS ||f - A psi|| / sqrt(N) (not in the reference; SURVEY F6) */
int mg_residual_norm(mg_ctx *ctx, double *rms);

/* ---- per-operator entry points on device pointers (rows a1-a7 of SURVEY section 8) */
int mg_twogrid(mg_ctx *ctx, double h, void *u, const void *f, int L);        /* cpu-raw.lua:186 */
int mg_smooth(mg_ctx *ctx, int L, void *u, const void *f, double h, int n);  /* n x cpu-raw.lua:176 */
int mg_jacobi(mg_ctx *ctx, int L, void *dest, const void *u, const void *f, double h);
int mg_residual(mg_ctx *ctx, int L, void *r, const void *f, const void *u, double h);
int mg_restrict(mg_ctx *ctx, int L2, void *R, const void *r);
int mg_prolong(mg_ctx *ctx, int L2, void *v, const void *V);
int mg_add_to(mg_ctx *ctx, size_t n, void *u, const void *v);
int mg_frob_err(mg_ctx *ctx, double *err);                   /* cpu-raw.lua:249-254 */
/* fused building blocks of MG_MODE_FUSED, exposed so each can be checked against the
 * composition of reference operators it replaces:
 *   mg_smooth_residual_restrict: n sweeps on u, then R = restrict(f - A u)
 *   mg_prolong_add_smooth:       u += prolong(V), then n sweeps on u                     */
int mg_smooth_residual_restrict(mg_ctx *ctx, int L, void *u, const void *f, double h, int n,
                                void *R);
int mg_prolong_add_smooth(mg_ctx *ctx, int L, void *u, const void *f, double h, int n,
                          const void *V);

/* ---- Krylov comparator of test/converge-multigrid-vs-krylov.lua:38-69 (SURVEY 8(f) rank 2).
 * Conjugate gradient on the reference's operator A(u) = (sum of neighbours - 4u)/h^2 (:48-58;
 * 3-D: six neighbours, -6u) with b = f and x = psi as found (the experiment starts from -f, which
 * is what initCells leaves in psi). err = ||r||_2/||b||_2 per iteration goes to err_hist, ||x||_inf
 * (what the experiment plots, :62-64) to linf_hist; stops at err < epsilon or max_iter. The
 * reference's `solver.conjgrad` is an un-vendored, unpinned dependency: parity unpinned.
 * mg_linf_norm: ||field||_inf, the quantity the experiment records per multigrid cycle (:25). */
int mg_cg(mg_ctx *ctx, int max_iter, double epsilon, double *err_hist, double *linf_hist, int *n_done);
int mg_linf_norm(mg_ctx *ctx, int which, int level, double *out);

/* ---- stage trace: the reference's `debugging` dumps (cpu-raw.lua:126-140) as records.
 * Only MG_MODE_REFSEQ records. name is one of 'f','u','r','R','V','v'. */
int mg_trace_enable(mg_ctx *ctx, int on);
int mg_trace_clear(mg_ctx *ctx);
size_t mg_trace_count(mg_ctx *ctx);
int mg_trace_get(mg_ctx *ctx, size_t i, char *name, int *L, const void **host_data, size_t *bytes);

/* ---- measurement helpers */
/* run n V-cycles back to back on the handle's stream between two CUDA events */
int mg_time_vcycles(mg_ctx *ctx, int n, float *ms_total);
/* kernels launched by this handle since creation (graph replays count their nodes) */
uint64_t mg_launch_count(mg_ctx *ctx);
/* one fused V-cycle without the graph, a CUDA-event pair around every launch; fills up to cap
 * records. kind: 0 smoother pass, 1 residual+restrict, 2 persistent small-level kernel,
 * 3 prolong+add, 4 copy, 5 smoother pass with fused prolong+add, 6 smoother pass with fused
 * residual+restrict. L = level width, sweeps = Jacobi sweeps done by that launch. */
int mg_profile_vcycle(mg_ctx *ctx, int cap, int *kind, int *L, int *sweeps, float *ms, int *n);

/* ---- multi-GPU slabs (3-D only; the reference has no multi-device path). The grid is cut along
 * z into nranks slabs; distributed levels carry ghost planes that are refreshed before every
 * smoother pass; coarse levels below the threshold are replicated. Results are bit-identical to
 * the single-GPU solver. Two transports:
 *  mg_create_slab       one handle per process/GPU, halo planes by ncclSend/ncclRecv (libnccl is
 *                       dlopen()ed). Rank 0 calls mg_nccl_unique_id and ships the 128 bytes to
 *                       the other ranks by any means. mg_upload/mg_download/mg_step_host then
 *                       move the planes THIS rank owns (size/nranks planes of size^2 elements).
 *  mg_create_slab_local every slab in this process on one device (testing the slab schedule on
 *                       a single GPU); the handle behaves like a single solver on the global grid. */
int mg_nccl_unique_id(void *id, size_t bytes);
int mg_create_slab(int dim, int size, int real_kind, int smooth, int device, int rank, int nranks,
                   const void *nccl_id, size_t id_bytes, mg_ctx **out);
int mg_create_slab_local(int dim, int size, int real_kind, int smooth, int device, int nslabs, mg_ctx **out);
/*  mg_create_slab_multi one process, one GPU per slab: what a single LuaJIT process (the reference's host,
 *                       test/test.lua:53-56; gpu.lua:27-30 picks ONE device) uses to drive 2, 4 or 8 GPUs. devices =
 *                       ndev device ordinals (NULL: 0 .. ndev-1), which must be able to access each other's memory.
 *                       The handle behaves like a single solver on the global grid (upload / download / step take the
 *                       whole field); kernels, fused halo stores and in-kernel handshakes are those of mg_create_slab. */
int mg_create_slab_multi(int dim, int size, int real_kind, int smooth, int ndev, const int *devices, mg_ctx **out);
/* process-wide defaults read by the next mg_create*: "slab_min_planes" (a level is cut across the ranks while every
 * rank keeps at least this many planes, default 32, >= 8; thinner levels are replicated -- the `cpuDepth` idea of
 * cpu-gpu.lua:11-15 at the scale of GPUs). The environment variable MGPOISSON_SLAB_MIN_PLANES still overrides it. */
int mg_set_global_option(const char *name, int value);
/* Fused halo exchange: every rank exports the CUDA IPC handle of its arena (64 bytes); after an
 * all-gather each rank attaches every other rank's arena. From then on the smoother kernel that
 * produces a boundary plane stores it straight into the neighbour's ghost planes over NVLink, the
 * kernel that restricts into the first replicated level stores into every rank's copy of it, and
 * the only inter-GPU operations of a V-cycle are flag handshakes inside those kernels: no NCCL call
 * is left in the cycle, which is replayed from one CUDA graph (option "slab_graph"). Option
 * "slab_p2p" = 0 goes back to ncclSend/ncclRecv/ncclAllGather. mg_create_slab_local and
 * mg_create_slab_multi use the same path with plain pointers. A rank whose neighbour stays silent
 * for 20 s gives up; the next synchronising call then returns MG_ESTATE. */
int mg_slab_ipc_export(mg_ctx *ctx, void *handle, size_t bytes);
int mg_slab_ipc_attach(mg_ctx *ctx, const void *handles, size_t bytes);
int mg_slab_info(mg_ctx *ctx, int *rank, int *nranks, int *own_planes, int *ghost, uint64_t *exchanges,
                 uint64_t *exchanged_bytes);
/* bytes stored straight into other GPUs' memory by this handle's kernels (fused halo exchange + fused all-gather),
 * and bytes moved by explicit exchanges (NCCL send/recv or peer copies), since creation */
int mg_slab_traffic(mg_ctx *ctx, uint64_t *peer_store_bytes, uint64_t *exchange_bytes);
/* measurement: after mg_set_option(ctx, "slab_trace", 1) every distributed smoother pass of this rank records 4 words
 * {ns on entry, ns after the wait for the lower neighbour, ns its first CTA waited for the upper neighbour, ns when its
 * last CTA finished} (this GPU's globaltimer). Copies up to cap records (the first ones since the option was set). */
int mg_slab_trace(mg_ctx *ctx, uint64_t *records, size_t cap, size_t *n);
]]

local lib = ffi.load(os.getenv'MGPOISSON_LIB' or 'mgpoisson')

local realKinds = {double = lib.MG_REAL_F64, float = lib.MG_REAL_F32, float_acc64 = lib.MG_REAL_F32_ACC64}
local ctypes = {double = 'double', float = 'float', float_acc64 = 'float'}

local function check(self, rc)
	if rc ~= 0 then
		error('libmgpoisson: '..ffi.string(lib.mg_last_error(self and self.handle or nil))..' ('..rc..')')
	end
end

local MultigridCUDA = class()

MultigridCUDA.debugging = false		-- cpu-raw.lua:121
MultigridCUDA.smooth = 7			-- cpu-raw.lua:123
MultigridCUDA.accuracy = 1e-10		-- cpu-raw.lua:124
MultigridCUDA.maxiter = 2			-- cpu-raw.lua:245 `for iter=1,2`
MultigridCUDA.dim = 2				-- 3 = this library's extension

-- level table: self.rs[L] etc. resolve to device pointers on demand (cpu-raw.lua:155-164)
local function levelTable(self, which)
	return setmetatable({}, {__index = function(_, L)
		return lib.mg_device_ptr(self.handle, which, L)
	end})
end

function MultigridCUDA:init(size, real, cpuDepth)
	if type(size) == 'table' then
		-- cpu.lua:173-181: MultigridCPU{size=, maxiter=, epsilon=, errorCallback=, debug=}, the form
		-- test/converge-multigrid-vs-krylov.lua:20-29 uses. cpu.lua starts the coarse corrections of
		-- every cycle from zero (cpu.lua:138), so this form does too.
		local args = size
		self.maxiter = args.maxiter or 1000			-- cpu.lua:22
		self.epsilon = args.epsilon or 1e-10		-- cpu.lua:21
		self.errorCallback = args.errorCallback
		if args.debug ~= nil then self.debugging = args.debug end
		self.zeroCorrections = true
		size, real, cpuDepth = args.size, args.real, args.cpuDepth
	end
	self.real = real or 'double'		-- cpu-raw.lua:143
	self.size = size
	local out = ffi.new'mg_ctx*[1]'
	check(nil, lib.mg_create(self.dim, size, assert(realKinds[self.real], 'unknown real'), self.smooth, -1, out))
	self.handle = ffi.gc(out[0], lib.mg_destroy)
	if cpuDepth then	-- cpu-gpu.lua:11-15: levels L <= 2^cpuDepth go to the small-level executor
		check(self, lib.mg_set_tuning(self.handle, -1, math.min(bit.lshift(1, cpuDepth), 256, size), -1))
	end
	if self.debugging then
		check(self, lib.mg_set_mode(self.handle, lib.MG_MODE_REFSEQ))
	end
	for name,which in pairs{f=lib.MG_BUF_F, psi=lib.MG_BUF_PSI, psiOld=lib.MG_BUF_PSIOLD,
							errorBuf=lib.MG_BUF_ERRORBUF, tmpU=lib.MG_BUF_TMPU} do
		local mg = self
		self[name] = {
			buffer = lib.mg_device_ptr(self.handle, which, size),
			-- mg.psi:normLInf(), what the convergence experiment records per cycle
			-- (converge-multigrid-vs-krylov.lua:25), computed on the device
			normLInf = function()
				local out = ffi.new'double[1]'
				check(mg, lib.mg_linf_norm(mg.handle, which, size, out))
				return out[0]
			end,
		}
	end
	self.rs = levelTable(self, lib.MG_BUF_r)
	self.Rs = levelTable(self, lib.MG_BUF_R)
	self.vs = levelTable(self, lib.MG_BUF_v)
	self.Vs = levelTable(self, lib.MG_BUF_V)
end

-- host <-> device, replacing enqueueReadBuffer/enqueueWriteBuffer (gpu.lua:263-267, cpu-gpu.lua:26-48)
function MultigridCUDA:getbuffer(which, L)
	local n = L^self.dim
	local cpuMem = ffi.new(ctypes[self.real]..'[?]', n)
	check(self, lib.mg_download(self.handle, which, L, cpuMem, n * ffi.sizeof(ctypes[self.real])))
	return cpuMem
end
function MultigridCUDA:setbuffer(which, L, cpuMem)
	check(self, lib.mg_upload(self.handle, which, L, cpuMem, L^self.dim * ffi.sizeof(ctypes[self.real])))
end

-- cpu-raw.lua:126-140: with `debugging` on, print a field the way the CPU versions do, so that runs can be diffed:
-- the name, then L rows of L values, each preceded by a blank (io.write(' ', im[j+L*i])); a non-finite value is an error.
-- `im` is HOST memory here (indexable from 0): a getbuffer() result or a stage record of the device trace.
-- dim = 3: L planes of that layout follow each other (the reference is 2-D only).
function MultigridCUDA:show(name, im, L)
	if not self.debugging then return end
	print(name)
	local rows = self.dim == 3 and L * L or L
	for i=0,rows-1 do
		for j=0,L-1 do
			io.write(' ', im[j+L*i])
		end
		print()
	end
	for i=0,rows*L-1 do
		if not math.isfinite(im[i]) then
			error("found a nan")
		end
	end
end

-- The reference's twoGrid calls show() between its operators (cpu-raw.lua:187-236). Here the whole cycle runs on the
-- device; with `debugging` the library runs one kernel per reference operator (MG_MODE_REFSEQ) and records every field
-- the reference would have shown, in its order -- replayed through show() afterwards.
function MultigridCUDA:showTrace()
	local n = tonumber(lib.mg_trace_count(self.handle))
	local name, L = ffi.new'char[1]', ffi.new'int[1]'
	local data, bytes = ffi.new'const void*[1]', ffi.new'size_t[1]'
	for i=0,n-1 do
		check(self, lib.mg_trace_get(self.handle, i, name, L, data, bytes))
		self:show(string.char(name[0]), ffi.cast(ctypes[self.real]..'*', data[0]), L[0])
	end
	check(self, lib.mg_trace_clear(self.handle))
end

-- conjugate gradient on the same operator, b = f, x = psi as found (converge-multigrid-vs-krylov.lua:38-69, whose
-- solver.conjgrad is an un-vendored library): returns the number of iterations and the per-iteration histories
-- err = |r|/|b| and |x|_inf (what the experiment's errorCallback records, :62-64)
function MultigridCUDA:conjgrad(maxiter, epsilon)
	maxiter = maxiter or 1000
	local errs, linf = ffi.new('double[?]', maxiter), ffi.new('double[?]', maxiter)
	local n = ffi.new'int[1]'
	check(self, lib.mg_cg(self.handle, maxiter, epsilon or 1e-20, errs, linf, n))
	return n[0], errs, linf
end

function MultigridCUDA:initCells()								-- cpu-raw.lua:8-20
	check(self, lib.mg_init_cells(self.handle))
end

function MultigridCUDA:inPlaceIterativeSolver(L, u, f, h)		-- cpu-raw.lua:176-184
	check(self, lib.mg_smooth(self.handle, L, u, f, h, 1))
end

function MultigridCUDA:twoGrid(h, u, f, L)						-- cpu-raw.lua:186-237
	if self.debugging then check(self, lib.mg_trace_enable(self.handle, 1)) end
	check(self, lib.mg_twogrid(self.handle, h, u, f, L))
	if self.debugging then self:showTrace() end
end

function MultigridCUDA:step()									-- cpu.lua:196-206
	if self.zeroCorrections then check(self, lib.mg_zero_corrections(self.handle)) end		-- cpu.lua:138
	local err = ffi.new'double[1]'
	if self.debugging then check(self, lib.mg_trace_enable(self.handle, 1)) end
	check(self, lib.mg_step(self.handle, err))
	if self.debugging then self:showTrace() end		-- the dumps of this cycle's twoGrid, in the reference's order
	return err[0]
end

function MultigridCUDA:run()									-- cpu-raw.lua:239-258
	print('#iter','err')
	for iter=1,self.maxiter do
		local err = self:step()
		print(iter, err)
		if err < self.accuracy or not math.isfinite(err) then break end
	end
end

-- cpu.lua:208-216 API shape (used by test/converge-multigrid-vs-krylov.lua:20-29)
function MultigridCUDA:solve()
	for iter=1,(self.maxiter or 1000) do
		local err = self:step()
		if self.errorCallback and self.errorCallback(iter, err) then break end
		if err < (self.epsilon or self.accuracy) or not math.isfinite(err) then break end
	end
end

return MultigridCUDA
