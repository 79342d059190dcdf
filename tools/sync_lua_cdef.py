"""Re-embeds the MGPOISSON_CDEF block of include/mgpoisson.h into the LuaJIT wrapper's ffi.cdef."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
hdr = open(os.path.join(ROOT, "include", "mgpoisson.h")).read()
cdef = re.search(r"/\* MGPOISSON_CDEF_BEGIN \*/\n(.*)/\* MGPOISSON_CDEF_END \*/", hdr, re.S).group(1)
path = os.path.join(ROOT, "lua-multigrid-poisson_b200", "lua", "multigrid-poisson", "cuda.lua")
lua = open(path).read()
new = re.sub(r"ffi\.cdef\[\[\n.*?\n\]\]", lambda m: "ffi.cdef[[\n" + cdef.rstrip("\n") + "\n]]", lua, count=1, flags=re.S)
if new != lua:
    open(path, "w").write(new)
    print("cuda.lua cdef updated")
