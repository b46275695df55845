"""Right-hand sides: from the reference's `funcptr` contract to device code.

The reference passes `funcptr = cfunc.address`, the address of x86 machine code
(README.md:24-33, numbalsoda/driver.py:37).  That code cannot run on the GPU, so
the same Python function is compiled a second time, by numba.cuda, into LTO-IR
and linked into the solver kernels (libb200ode: b200ode_link_rhs).  Accepted
spellings of the right-hand side:

  * the integer address of a numba `@cfunc(lsoda_sig)` (what the reference takes):
    the CFunc object is found again among the live objects and its Python function
    recompiled;
  * the CFunc object itself, or the plain Python function `rhs(t, u, du, p)`;
  * `builtin("lorenz")` etc.: device code shipped inside libb200ode.
"""
import ctypes as ct
import gc
import numbers
import threading
import weakref

from . import _lib

_lock = threading.Lock()
_handles = {}     # (key, solver, neq, flags) -> RhsHandle
_by_address = {}  # cfunc address -> weakref to the CFunc (validated on every lookup)


class Builtin:
    def __init__(self, name):
        self.name = name

    def __repr__(self):
        return "builtin(%r)" % self.name


def builtin(name):
    """A right-hand side compiled into libb200ode (lorenz, robertson, lotka_volterra,
    brusselator1d)."""
    return Builtin(name)


class RhsHandle:
    """A linked (solver kernel + right-hand side) module; owns the C handle."""

    def __init__(self, ptr, solver, neq, flags, origin):
        self.ptr, self.solver, self.neq, self.flags, self.origin = ptr, solver, neq, flags, origin

    def cubin(self):
        p, n = ct.c_void_p(), ct.c_size_t()
        _lib.check(_lib.lib.b200ode_rhs_cubin(self.ptr, ct.byref(p), ct.byref(n)))
        return ct.string_at(p, n.value)

    def kernel_info(self):
        info = (ct.c_int * 8)()
        _lib.check(_lib.lib.b200ode_kernel_info(self.ptr, ct.byref(info)))
        keys = ["registers", "local_bytes", "static_smem", "threads_per_block", "blocks_per_sm",
                "sm_count", "last_grid", "launches"]
        return dict(zip(keys, list(info)))

    def __del__(self):
        try:
            if self.ptr:
                _lib.lib.b200ode_rhs_free(self.ptr)
                self.ptr = None
        except Exception:
            pass


def register(cfunc):
    """Remember a numba CFunc so that its address resolves without a scan of the live
    objects (optional)."""
    _by_address[cfunc.address] = weakref.ref(cfunc)
    return cfunc


def _cached(addr):
    """The Python function of the live CFunc at `addr`, or None.  A cache entry is only a
    hint: once a cfunc is collected numba may hand its code address to a new one, so the
    entry counts only while its CFunc is alive and still has that address."""
    ref = _by_address.get(addr)
    obj = ref() if ref is not None else None
    if obj is not None and obj.address == addr:
        return obj._pyfunc
    _by_address.pop(addr, None)
    return None


def _pyfunc_from_address(addr):
    addr = int(addr)
    f = _cached(addr)
    if f is not None:
        return f
    from numba.core.ccallback import CFunc
    for obj in gc.get_objects():
        try:  # (isinstance itself raises on dead weakref proxies)
            if isinstance(obj, CFunc):
                _by_address[obj.address] = weakref.ref(obj)
        except Exception:
            pass
    f = _cached(addr)
    if f is not None:
        return f
    raise ValueError(
        "funcptr %#x is not the address of a live numba @cfunc: the GPU solver needs the "
        "Python source of the right-hand side to compile it for the device; pass the cfunc "
        "object or the Python function instead" % addr)


def without_carray(pyfunc, record_data=False):
    """The reference's right-hand sides view their raw pointers with `nb.carray`
    (benchmark/benchmark_lorenz.py:10-13; passing_data_to_rhs_function.ipynb); numba.cuda cannot
    compile carray.  Where the view is only a change of type the source is rewritten:

      x = nb.carray(ptr, (n,))          ->  x = ptr               (1-D view of a double pointer)
      a, b, c = x                       ->  a = x[0]; b = x[1]; c = x[2]
      nb.carray(address_as_void_pointer(a), shape, dtype=T)
                                        ->  address_as_pointer(a, T)    (indexable typed pointer)
    and, when the data argument is a record (record_data: the 4th parameter is then typed as
    the record itself, which numba passes by reference):
      d = nb.carray(data_p, 1); d[0].f  ->  data_p.f
      nb.carray(data_p, 1)[0].f         ->  data_p.f
    CPU dispatchers (`nb.njit` functions) the body calls are compiled as device functions.  The
    arithmetic is untouched."""
    import ast
    import inspect
    import textwrap
    tree = ast.parse(textwrap.dedent(inspect.getsource(pyfunc)))
    fdef = tree.body[0]
    if not isinstance(fdef, ast.FunctionDef):
        raise TypeError("cannot rewrite %r" % (pyfunc,))
    fdef.decorator_list = []
    data_param = fdef.args.args[3].arg if len(fdef.args.args) >= 4 else None
    views = {}
    recviews = {}

    def fname(node):
        f = node.func if isinstance(node, ast.Call) else None
        return getattr(f, "attr", None) or getattr(f, "id", None)

    def is_carray(node):
        return fname(node) == "carray"

    def shape_is_one(node):
        shape = node.args[1] if len(node.args) > 1 else None
        if isinstance(shape, ast.Tuple) and len(shape.elts) == 1:
            shape = shape.elts[0]
        return isinstance(shape, ast.Constant) and shape.value == 1

    def is_record_view(node):
        return (record_data and is_carray(node) and isinstance(node.args[0], ast.Name) and
                node.args[0].id == data_param and shape_is_one(node))

    class Rewrite(ast.NodeTransformer):
        def visit_Call(self, node):
            self.generic_visit(node)
            if is_carray(node) and node.args and fname(node.args[0]) == "address_as_void_pointer":
                dt = node.args[2] if len(node.args) > 2 else None
                for kw in node.keywords:
                    if kw.arg == "dtype":
                        dt = kw.value
                if dt is None:
                    raise TypeError("nb.carray on a void pointer needs dtype=")
                return ast.copy_location(ast.Call(func=ast.Name(id="__b200_address_as_pointer", ctx=ast.Load()),
                                                  args=[node.args[0].args[0], dt], keywords=[]), node)
            return node

        def visit_Subscript(self, node):
            self.generic_visit(node)
            idx = node.slice
            zero = isinstance(idx, ast.Constant) and idx.value == 0
            if zero and is_record_view(node.value):
                return ast.copy_location(ast.Name(id=data_param, ctx=ast.Load()), node)
            if zero and isinstance(node.value, ast.Name) and node.value.id in recviews:
                return ast.copy_location(ast.Name(id=data_param, ctx=ast.Load()), node)
            return node

        def visit_Assign(self, node):
            t = node.targets[0] if len(node.targets) == 1 else None
            if isinstance(t, ast.Name) and is_record_view(node.value):
                recviews[t.id] = True
                return None  # the view itself disappears; its uses become the record parameter
            self.generic_visit(node)
            if isinstance(t, ast.Name) and is_carray(node.value):
                shape = node.value.args[1] if len(node.value.args) > 1 else None
                dims = shape.elts if isinstance(shape, ast.Tuple) else [shape]
                if len(dims) != 1:
                    raise TypeError("only 1-D nb.carray views can run on the device")
                views[t.id] = True
                return ast.copy_location(ast.Assign(targets=[t], value=node.value.args[0]), node)
            if isinstance(t, ast.Tuple) and isinstance(node.value, ast.Name) and node.value.id in views:
                out = []
                for k, e in enumerate(t.elts):
                    sub = ast.Subscript(value=ast.Name(id=node.value.id, ctx=ast.Load()),
                                        slice=ast.Constant(k), ctx=ast.Load())
                    out.append(ast.copy_location(ast.Assign(targets=[e], value=sub), node))
                return out
            return node

    tree = ast.fix_missing_locations(Rewrite().visit(tree))
    env = dict(pyfunc.__globals__)
    cv = inspect.getclosurevars(pyfunc)
    env.update(cv.nonlocals)
    from .driver import address_as_pointer
    env["__b200_address_as_pointer"] = address_as_pointer
    # nb.njit functions called from the body: the same Python function as a device function
    try:
        from numba import cuda
        from numba.core.dispatcher import Dispatcher
        for k, val in list(env.items()):
            if isinstance(val, Dispatcher) and getattr(val, "py_func", None) is not None:
                env[k] = cuda.jit(device=True)(val.py_func)
    except Exception:
        pass
    exec(compile(tree, "<numbalsoda_b200: %s without carray>" % pyfunc.__name__, "exec"), env)
    return env[fdef.name]


_TARGET_CC = (10, 0)  # the solver images are built for sm_100a (build.py ARCH)


def _ltoir_cc():
    """Compute capability handed to numba's NVVM for the right-hand side's LTO-IR.  LTO-IR is
    target-independent up to the NVVM data layout; nvJitLink generates the sm_100a code.  The
    target is the solver images' (10.0), lowered to the newest architecture the installed NVVM
    knows when it predates Blackwell."""
    try:
        from numba.cuda.cudadrv import nvvm
        ccs = [tuple(c[:2]) for c in nvvm.get_supported_ccs()]
        ok = [c for c in ccs if c <= _TARGET_CC]
        if ok:
            return max(ok)
    except Exception:
        pass
    return (9, 0)


class _nested_device_functions_for:
    """Device functions the right-hand side calls (cuda.jit(device=True) dispatchers, e.g. the
    nb.njit helpers without_carray converts) are compiled by numba for the *current device's*
    compute capability, which needs a CUDA context.  The image is LTO-IR for nvJitLink, linked
    for sm_100a whatever numba was told, and linking must work on machines without a GPU: for
    the duration of a compile the nested dispatchers are told the compute capability the outer
    function is compiled for."""

    class _Device:
        def __init__(self, cc):
            self.compute_capability = cc

    def __init__(self, cc):
        self.cc = cc

    def __enter__(self):
        import numba.cuda.dispatcher as disp
        self.disp, self.saved = disp, disp.get_current_device
        cc = self.cc
        disp.get_current_device = lambda: _nested_device_functions_for._Device(cc)

    def __exit__(self, *exc):
        self.disp.get_current_device = self.saved
        return False


def compile_ltoir(pyfunc, abi_name="b200ode_user_rhs", data_dtype=None, lanes=False):
    """numba.cuda: Python rhs(t, u, du, p) -> LTO-IR of `b200ode_user_rhs` (or, for
    solve_ivp's events(t, y, out, p), of `b200ode_user_events`).  data_dtype: a structured numpy
    dtype -- `data` is then a record and the function receives it as one (numba passes records
    by reference, so the C ABI still sees the record's address)."""
    from numba import cuda, types
    ptype = types.CPointer(types.double)
    if data_dtype is not None:
        import numba
        ptype = numba.from_dtype(data_dtype)
    sig = (types.double, types.CPointer(types.double), types.CPointer(types.double), ptype)
    if lanes:  # rhs_lanes(t, u, du, p, lane, nlanes): the lane-parallel form (include/b200ode.h)
        sig = sig + (types.int32, types.int32)

    def build(f):
        with _nested_device_functions_for(_ltoir_cc()):
            code, _ = cuda.compile(f, sig, device=True, abi="c",
                                   abi_info={"abi_name": abi_name}, output="ltoir", cc=_ltoir_cc())
        return bytes(code)

    try:
        return build(pyfunc)
    except Exception as first:
        try:
            import inspect
            if "carray" not in inspect.getsource(pyfunc):
                raise first
            rewritten = without_carray(pyfunc, record_data=data_dtype is not None)
        except Exception:
            raise first
        return build(rewritten)


def _is_address(x):
    """funcptr as the reference takes it: a Python int or any numpy integer scalar."""
    return isinstance(x, numbers.Integral) and not isinstance(x, bool)


def get_handle(rhs, solver, neq, strict=False, exact_ctrl=False, data_dtype=None, rhs_lanes=None):
    """Link (once per right-hand side, solver, size, mode and record type of `data`) and return
    the handle."""
    flags = _lib.STRICT_FP if strict else (_lib.EXACT_CTRL if exact_ctrl else 0)
    if isinstance(rhs, RhsHandle):
        return rhs
    if isinstance(rhs, Builtin):
        key, origin = ("builtin", rhs.name), rhs
    else:
        if _is_address(rhs):
            rhs = _pyfunc_from_address(rhs)
        elif hasattr(rhs, "_pyfunc") and hasattr(rhs, "address"):
            rhs = rhs._pyfunc
        if not callable(rhs):
            raise TypeError("right-hand side must be a cfunc address, a cfunc, a Python function "
                            "or builtin(name); got %r" % (rhs,))
        key, origin = ("py", rhs), rhs
    if rhs_lanes is not None and hasattr(rhs_lanes, "_pyfunc"):
        rhs_lanes = rhs_lanes._pyfunc
    k = (key, solver, neq, flags, str(data_dtype) if data_dtype is not None else None, rhs_lanes)
    with _lock:
        h = _handles.get(k)
        if h is not None:
            return h
        out = ct.c_void_p()
        if isinstance(origin, Builtin):
            name = origin.name.encode()
            rc = _lib.lib.b200ode_link_rhs(solver, neq, name, len(name), _lib.IMG_BUILTIN, flags,
                                           ct.byref(out))
        else:
            img = compile_ltoir(origin, data_dtype=data_dtype)
            if rhs_lanes is not None and neq > 8:
                limg = compile_ltoir(rhs_lanes, "b200ode_user_rhs_lanes", data_dtype=data_dtype, lanes=True)
                rc = _lib.lib.b200ode_link_rhs_lanes(solver, neq, img, len(img), _lib.IMG_LTOIR, limg, len(limg),
                                                     _lib.IMG_LTOIR, flags, ct.byref(out))
            else:
                rc = _lib.lib.b200ode_link_rhs(solver, neq, img, len(img), _lib.IMG_LTOIR, flags,
                                               ct.byref(out))
        _lib.check(rc)
        h = RhsHandle(out, solver, neq, flags, origin)
        _handles[k] = h
        return h


def _origin(f, what):
    """-> (cache key, Builtin or Python function) of a right-hand side / events function."""
    if isinstance(f, Builtin):
        return ("builtin", f.name), f
    if _is_address(f):
        f = _pyfunc_from_address(f)
    elif hasattr(f, "_pyfunc") and hasattr(f, "address"):
        f = f._pyfunc
    if not callable(f):
        raise TypeError("%s must be a cfunc address, a cfunc, a Python function or builtin(name); "
                        "got %r" % (what, f))
    return ("py", f), f


def get_ivp_handle(rhs, neq, n_events=0, events=None, strict=False, exact_ctrl=False):
    """solve_ivp: link the right-hand side and the events function (reference
    numbalsoda/driver_solve_ivp.py:86-88: `funcptr`, `n_events`, `event_fcn`) with the
    solve_ivp kernel, once per combination."""
    flags = _lib.STRICT_FP if strict else (_lib.EXACT_CTRL if exact_ctrl else 0)
    if isinstance(rhs, RhsHandle):
        return rhs
    n_events = int(n_events)
    if n_events > 0 and (events is None or (_is_address(events) and int(events) == 0)):
        raise ValueError("n_events = %d but no event function was given" % n_events)
    rkey, rorig = _origin(rhs, "right-hand side")
    ekey, eorig = _origin(events, "event_fcn") if n_events > 0 else (None, None)
    k = (rkey, ekey, n_events, _lib.SOLVE_IVP, neq, flags)
    with _lock:
        h = _handles.get(k)
        if h is not None:
            return h

        def image(orig, abi_name):
            if isinstance(orig, Builtin):
                name = orig.name.encode()
                return name, len(name), _lib.IMG_BUILTIN
            img = compile_ltoir(orig, abi_name)
            return img, len(img), _lib.IMG_LTOIR
        ri, rn, rk = image(rorig, "b200ode_user_rhs")
        ei, en, ek = image(eorig, "b200ode_user_events") if n_events > 0 else (None, 0, 0)
        out = ct.c_void_p()
        _lib.check(_lib.lib.b200ode_link_ivp(neq, ri, rn, rk, n_events, ei, en, ek, flags, ct.byref(out)))
        h = RhsHandle(out, _lib.SOLVE_IVP, neq, flags, (rorig, eorig))
        _handles[k] = h
        return h
