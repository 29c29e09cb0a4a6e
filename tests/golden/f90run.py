"""A small Fortran-90-subset interpreter: runs subroutines of the REFERENCE'S OWN SOURCE TEXT in fp32.

Why.  The reference (TRISTAN-MP, Fortran + MPI) cannot be compiled in the build image or on the GPU box (no Fortran compiler:
profiles/r2_toolchain_probe.txt), so the CPU oracle (oracle/pic_oracle.c) -- a hand restatement of the hot-path routines --
had never been checked against anything produced by the reference.  This module reads a subroutine straight out of
/root/reference/code/*.F90, applies the cpp conditionals, translates the statements one by one into Python and executes them
with numpy.float32 scalars (every operation rounds to fp32, in the source's order; integer division truncates; arrays are
1-based, column-major, and -- as the reference relies on -- not bounds-checked on the first index).  tests/golden/
make_ref_golden.py uses it to produce golden vectors (committed as tests/golden/ref_*.npz) that pin the oracle; nothing at test
time or on the GPU box reads /root/reference.

Subset: subroutines and functions (integer / real result), scalar / array assignments, whole-array assignments and
expressions, array sections, do (local or module loop variable) / do while / if-elseif-else, one-line if, the two goto idioms
of the path (an unconditional forward skip to `N continue`; `go to N` where `N continue` closes the loop = cycle), call (to
other translated routines or to Python stand-ins), derived-type components (p(n)%x) and element assignment (tempp(j) = p(n)),
integer arrays, real(dprec) / double precision entities and d-exponent literals, DATA, optional arguments with present(),
select case on strings, where / elsewhere, cycle / exit, allocate, unformatted sequential READ / WRITE with implied DO (records
framed as gfortran frames them), the intrinsics aint int real dble min max abs sqrt sum cshift mod
dmod modulo sign ceiling nint floor and exp log log10 cos sin atan tan ** (these through glibc's libm, float or double entry
point by operand kind).  Scalar actual arguments of a `call` receive the callee's final dummy values (by-reference semantics);
functions that update an argument inside an expression (random(dseed)) work through `alias_globals`.

MPI.  MPI_SendRecv is executed, not stubbed: `count` elements of the send buffer in Fortran element order (Payload) go to
`dest` with `sendtag`, and the message from `source` with `recvtag` lands in the first `count` elements of the receive buffer.
With one rank (Globals.comm is None) every neighbour is the rank itself.  With several ranks each rank is a THREAD that runs
the reference's text on its own Globals, and Comm.sendrecv is the rendezvous -- neighbour ranks, tags, counts and ordering
are whatever the reference's source computes.  Other mpi_* calls (barrier, wtime as a stand-in) carry no data and are dropped.
TEST INFRASTRUCTURE ONLY.
"""
import math
import re

import numpy as np

F = np.float32


# ------------------------------------------------------------------------------------------------------------------
# run-time support
# ------------------------------------------------------------------------------------------------------------------
class FArr:
    """1-based, column-major Fortran array.  a(i,j,k) with an out-of-range first index addresses the flat storage, as the
    reference's curx(l2,1,1) trick does (particles.F90:1058)."""

    __array_ufunc__ = None               # numpy scalars defer to the reflected operators below

    def __init__(self, shape, dtype=np.float32, data=None):
        shape = tuple(int(s) for s in (shape if isinstance(shape, (tuple, list)) else (shape,)))
        self.shape = shape
        self.flat = np.zeros(int(np.prod(shape)), dtype) if data is None else data
        self.strides = [1]
        for s in shape[:-1]:
            self.strides.append(self.strides[-1] * s)

    @classmethod
    def from_c(cls, a):
        """from a C-ordered numpy array shaped (mz, my, mx) == Fortran (mx, my, mz); shares memory"""
        shape = tuple(reversed(a.shape))
        return cls(shape, a.dtype, a.reshape(-1))

    def nd(self):
        return self.flat.reshape(self.shape, order="F")

    def _scalar_index(self, key):
        if not isinstance(key, tuple):
            key = (key,)
        off = 0
        for k, st in zip(key, self.strides):
            off += (int(k) - 1) * st
        if off < 0 or off >= self.flat.size:
            raise IndexError(f"flat index {off} outside the array ({self.flat.size})")
        return off

    @staticmethod
    def _is_scalar_key(key):
        key = key if isinstance(key, tuple) else (key,)
        return all(not isinstance(k, slice) for k in key)

    def _section(self, key):
        key = key if isinstance(key, tuple) else (key,)
        idx = []
        for k, n in zip(key, self.shape):
            if isinstance(k, slice):
                lo = 1 if k.start is None else int(k.start)
                hi = n if k.stop is None else int(k.stop)
                idx.append(slice(lo - 1, hi, None if k.step is None else int(k.step)))
            else:
                idx.append(int(k) - 1)
        return tuple(idx)

    def __getitem__(self, key):
        if self._is_scalar_key(key):
            v = self.flat[self._scalar_index(key)]
            return int(v) if self.flat.dtype.kind == "i" else v       # Fortran integers stay integers in mixed expressions
        return np.array(self.nd()[self._section(key)])                # sections are values

    def __setitem__(self, key, val):
        if self._is_scalar_key(key):
            self.flat[self._scalar_index(key)] = val
        elif isinstance(val, Payload):
            dst = self.nd()[self._section(key)]
            tmp = np.array(dst).reshape(-1, order="F")
            tmp[:val.data.size] = val.data
            dst[...] = tmp.reshape(dst.shape, order="F")
        else:
            v = val.nd() if isinstance(val, FArr) else val
            dst = self.nd()[self._section(key)]
            if isinstance(v, np.ndarray) and v.shape != dst.shape and v.ndim == dst.ndim:
                # non-conforming section assignment (quirk Q6, fieldboundaries.F90:1938: g layers on the left, g + 1 on the
                # right): compiled code loops over the LEFT side's extents; the surplus on the right is not copied
                sl = tuple(slice(0, min(a, b)) for a, b in zip(dst.shape, v.shape))
                dst[sl] = v[sl]
            else:
                dst[...] = v

    def set(self, val):
        """whole-array assignment"""
        if isinstance(val, Payload):
            if isinstance(val.rtype, Subarray):
                if self.shape != val.rtype.sizes:
                    raise TypeError(f"subarray datatype {val.rtype.sizes} on a receive buffer of shape {self.shape}")
                dst = self.nd()[val.rtype.box]
                dst[...] = val.data.reshape(dst.shape, order="F")
                return
            self.flat[:val.data.size] = val.data
            return
        self.nd()[...] = val.nd() if isinstance(val, FArr) else val

    # whole-array arithmetic (elementwise, fp32)
    def _bin(self, other, op, rev=False):
        o = other.nd() if isinstance(other, FArr) else other
        a = self.nd()
        with np.errstate(all="ignore"):
            r = op(o, a) if rev else op(a, o)
        out = FArr(self.shape, self.flat.dtype)
        out.nd()[...] = r
        return out

    def __add__(self, o): return self._bin(o, np.add)
    def __radd__(self, o): return self._bin(o, np.add, True)
    def __sub__(self, o): return self._bin(o, np.subtract)
    def __rsub__(self, o): return self._bin(o, np.subtract, True)
    def __mul__(self, o): return self._bin(o, np.multiply)
    def __rmul__(self, o): return self._bin(o, np.multiply, True)
    def __truediv__(self, o): return self._bin(o, np.divide)
    def __neg__(self): return self._bin(F(-1), np.multiply)
    def __ne__(self, o): return self.nd() != (o.nd() if isinstance(o, FArr) else o)      # masks of WHERE constructs
    def __eq__(self, o): return self.nd() == (o.nd() if isinstance(o, FArr) else o)
    def __gt__(self, o): return self.nd() > (o.nd() if isinstance(o, FArr) else o)
    def __lt__(self, o): return self.nd() < (o.nd() if isinstance(o, FArr) else o)
    __hash__ = None

    def where_set(self, mask, val):
        """a = val inside WHERE (mask)"""
        with np.errstate(all="ignore"):
            v = val.nd() if isinstance(val, FArr) else val
            self.nd()[mask] = v[mask] if isinstance(v, np.ndarray) else v


class Payload:
    """what an MPI message carries: `count` elements of the send buffer in Fortran (column-major) element order.  The receive
    side stores them into the first `count` elements of ITS buffer, again in Fortran order -- shapes play no role."""

    def __init__(self, data, rtype=None):
        self.data = data                  # 1-D numpy array (plain or structured), already a copy
        self.rtype = rtype                # receive-side derived datatype (Subarray) or None


class Subarray:
    """MPI_Type_create_subarray(ndims, sizes, subsizes, starts, MPI_ORDER_FORTRAN, ...): a box inside an array"""

    def __init__(self, sizes, subsizes, starts):
        self.sizes = tuple(int(v) for v in sizes)
        self.box = tuple(slice(int(a), int(a) + int(n)) for a, n in zip(starts, subsizes))


def as_payload(buf, count, stype=None):
    if isinstance(buf, Payload):
        return buf
    if isinstance(stype, Subarray):
        if not isinstance(buf, FArr) or buf.shape != stype.sizes or int(count) != 1:
            raise TypeError(f"subarray datatype {stype.sizes} on a buffer of shape {getattr(buf, 'shape', None)}, count {count}")
        return Payload(buf.nd()[stype.box].reshape(-1, order="F").copy())
    if isinstance(buf, RecArr):
        flat = buf.a
    elif isinstance(buf, FArr):
        flat = buf.flat
    else:
        flat = np.asarray(buf).reshape(-1, order="F")
    n = int(count)
    if n > flat.size:
        # The reference does this once: the 3D x-fold of exchange_current sizes its second message with the x extent instead
        # of the y extent (fieldboundaries.F90:1838), so with mx > my the count exceeds the section.  A compiled run sends
        # the section's contiguous temporary plus whatever follows it on the heap and the receiver copies back only the
        # section: the section's own elements arrive intact.  Modelled as "send what exists"; noted in OVERLONG.
        if isinstance(buf, (RecArr, FArr)):
            raise IndexError(f"MPI send of {n} elements from a buffer of {flat.size}")
        OVERLONG.append((n, flat.size))
        n = flat.size
    return Payload(flat[:n].copy())


OVERLONG = []


class Comm:
    """MPI_COMM_WORLD for ranks that run as Python threads: MPI_SendRecv = post the send, then block on the receive"""

    def __init__(self, size, timeout=30.0):
        import collections
        import queue
        self.size, self.timeout = size, timeout
        self.q = collections.defaultdict(queue.Queue)

    def allreduce(self, rank, val, op):
        """all ranks contribute, everybody gets the rank-ordered combination (op: 'sum' | 'max' | 'min')"""
        import threading
        if not hasattr(self, "_ar"):
            self._ar = {"lock": threading.Lock(), "vals": {}, "gen": 0, "cv": None, "res": None}
            self._ar["cv"] = threading.Condition(self._ar["lock"])
        st = self._ar
        with st["cv"]:
            gen = st["gen"]
            st["vals"][rank] = val
            if len(st["vals"]) == self.size:
                vs = [st["vals"][r] for r in range(self.size)]
                acc = vs if op == "gather" else vs[0]
                for v in ([] if op == "gather" else vs[1:]):
                    acc = (acc + v) if op == "sum" else (np.maximum(acc, v) if op == "max" else np.minimum(acc, v))
                st["res"], st["vals"], st["gen"] = acc, {}, gen + 1
                st["cv"].notify_all()
            else:
                if not st["cv"].wait_for(lambda: st["gen"] != gen, timeout=self.timeout):
                    raise TimeoutError("allreduce: a rank did not arrive")
            return st["res"]

    def sendrecv(self, rank, payload, dest, sendtag, source, recvtag):
        self.q[(rank, int(dest), int(sendtag))].put(payload)
        return self.q[(int(source), rank, int(recvtag))].get(timeout=self.timeout)


def run_ranks(fns):
    """run one callable per rank concurrently (they meet inside Comm.sendrecv); re-raises the first failure"""
    import threading
    err = []

    def wrap(f):
        try:
            f()
        except BaseException as e:        # noqa: BLE001 -- reported to the caller below
            err.append(e)
    th = [threading.Thread(target=wrap, args=(f,)) for f in fns]
    for t in th:
        t.start()
    for t in th:
        t.join()
    if err:
        raise err[0]


class RecordReader:
    """READ(unit) io-list on one unformatted record: items are taken in order, each in the kind of its target; reading past
    the end of the record is the run-time error it is in Fortran, reading less than the record holds is allowed"""
    KIND = {"int": ("<i4", int), "real": ("<f4", F), "real8": ("<f8", np.float64), "logical": ("<i4", bool)}

    def __init__(self, rec):
        n = int(np.frombuffer(rec[:4], "<i4")[0])
        if len(rec) != n + 8 or rec[-4:] != rec[:4]:
            raise IOError("bad record markers")
        self.b, self.pos = rec[4:-4], 0

    def take(self, dt):
        dt = np.dtype(dt)
        if self.pos + dt.itemsize > len(self.b):
            raise IOError(f"end of record: wanted {dt.itemsize} bytes at offset {self.pos} of {len(self.b)}")
        v = np.frombuffer(self.b, dt, 1, self.pos)[0]
        self.pos += dt.itemsize
        return v

    def scalar(self, kind):
        dt, cast = self.KIND[kind]
        return cast(self.take(dt))

    def elem(self, arr, idx):
        arr[idx if len(idx) > 1 else idx[0]] = self.take(arr.flat.dtype.newbyteorder("<"))

    def comp(self, recarr, i, name):
        recarr.a[name][int(i) - 1] = self.take(recarr.a.dtype[name].newbyteorder("<"))

    def close(self):
        self.left = len(self.b) - self.pos


class Record:
    """one element of an array of derived type (type particle)"""

    def __init__(self, arr, i):
        object.__setattr__(self, "_a", arr)
        object.__setattr__(self, "_i", i)

    def __getattr__(self, name):
        return self._a[name][self._i]

    def __setattr__(self, name, val):
        self._a[name][self._i] = val


class RecArr:
    """1-based array of derived type over a numpy structured array"""

    def __init__(self, a):
        self.a = a

    def __getitem__(self, i):
        return Record(self.a, int(i) - 1)

    def __setitem__(self, i, rec):
        self.a[int(i) - 1] = rec._a[rec._i]               # tempp(j) = p(n)

    def set(self, val):
        """whole-array assignment from a message or from another array of the type"""
        v = val.data if isinstance(val, Payload) else val.a
        self.a[:v.size] = v


def fdiv(a, b):
    if isinstance(a, (int, np.integer)) and isinstance(b, (int, np.integer)):
        return int(math.trunc(int(a) / int(b))) if abs(int(a)) < 2 ** 52 else int(a) // int(b)
    return a / b


def fpow(a, b):
    if isinstance(b, (int, np.integer)) and not isinstance(a, (int, np.integer)):
        # x**n with an integer n: multiplications by repeated squaring, as gcc's __builtin_powi and every compiler's
        # expansion of x**2, x**3 do (x**2 = x*x, x**3 = x*(x*x), x**4 = (x*x)*(x*x))
        K = (lambda v: v) if isinstance(a, (FArr, np.ndarray)) else (np.float64 if isinstance(a, np.float64) else F)
        n, x = abs(int(b)), a
        r = x if n & 1 else (np.float64(1.0) if isinstance(a, np.float64) else F(1.0))
        n >>= 1
        while n:
            x = K(x * x)
            if n & 1:
                r = K(r * x)
            n >>= 1
        return r if b >= 0 else K((np.float64(1.0) if isinstance(a, np.float64) else F(1.0)) / r)
    if isinstance(a, (int, np.integer)) and isinstance(b, (int, np.integer)):
        return int(a) ** int(b)
    if isinstance(a, np.float64) or isinstance(b, np.float64):
        return np.float64(_libm().pow(float(a), float(b)))
    return F(_libm().powf(float(F(a)), float(F(b))))                  # real ** real: libm's powf, as a compiled build calls


def fsum(a):
    """sum(): sequential, in fp32"""
    v = a.nd().reshape(-1, order="F") if isinstance(a, FArr) else np.asarray(a).reshape(-1, order="F")
    if v.dtype.kind == "i":
        return int(v.sum())
    K = np.float64 if v.dtype == np.float64 else F
    s = K(0.0)
    for x in v:
        s = K(s + x)
    return s


def fcshift(a, shift, dim=1):
    """cshift(array, shift, dim): result(i) = array(i + shift), circular"""
    out = FArr(a.shape, a.flat.dtype)
    out.nd()[...] = np.roll(a.nd(), -int(shift), axis=int(dim) - 1)
    return out


def frange(a, b, c=1):
    a, b, c = int(a), int(b), int(c)
    return range(a, b + (1 if c > 0 else -1), c)


def fexit(a, b, c=1):
    """value of a DO variable after normal termination"""
    a, b, c = int(a), int(b), int(c)
    n = max((b - a + c) // c, 0)
    return a + n * c


def faint(x):
    return np.trunc(x) if isinstance(x, np.float64) else F(np.trunc(F(x)))


def fmodulo(a, b):
    if isinstance(a, (int, np.integer)) and isinstance(b, (int, np.integer)):
        return int(a) - int(b) * math.floor(int(a) / int(b))
    r = a - b * np.floor(a / b)
    return r if isinstance(r, np.float64) else F(r)


def fmod(a, b):
    if isinstance(a, (int, np.integer)) and isinstance(b, (int, np.integer)):
        return int(math.fmod(int(a), int(b)))
    r = np.fmod(a, b)
    return r if isinstance(r, np.float64) else F(r)


def fmin(*a):
    return min(a)


def fmax(*a):
    return max(a)


_LIBM = None


def _libm():
    """glibc's libm: a compiled reference calls its float / double entry points (cosf, expf, log10f ...), and so does the
    oracle; numpy's own SIMD versions can differ from them in the last bit"""
    global _LIBM
    if _LIBM is None:
        import ctypes
        import ctypes.util
        _LIBM = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
        for nm in ("cos", "sin", "exp", "log", "log10", "atan", "tan"):
            getattr(_LIBM, nm).restype = ctypes.c_double
            getattr(_LIBM, nm).argtypes = [ctypes.c_double]
            getattr(_LIBM, nm + "f").restype = ctypes.c_float
            getattr(_LIBM, nm + "f").argtypes = [ctypes.c_float]
        _LIBM.pow.restype, _LIBM.pow.argtypes = ctypes.c_double, [ctypes.c_double, ctypes.c_double]
        _LIBM.powf.restype, _LIBM.powf.argtypes = ctypes.c_float, [ctypes.c_float, ctypes.c_float]
    return _LIBM


def elementary(name):
    def f(x):
        if isinstance(x, FArr):
            out = FArr(x.shape, x.flat.dtype)
            out.flat[:] = [f(v) for v in x.flat]
            return out
        if isinstance(x, np.float64):
            return np.float64(getattr(_libm(), name)(float(x)))
        return F(getattr(_libm(), name + "f")(float(F(x))))
    return f


def fsqrt(x):
    """IEEE square root in the operand's kind (correctly rounded everywhere)"""
    if isinstance(x, FArr):
        out = FArr(x.shape, x.flat.dtype)
        with np.errstate(all="ignore"):
            out.flat[:] = np.sqrt(x.flat)
        return out
    with np.errstate(all="ignore"):
        return np.sqrt(x) if isinstance(x, np.float64) else F(np.sqrt(F(x)))


def freal(x, kind=4):
    if isinstance(x, FArr):
        out = FArr(x.shape, np.float64 if int(kind) == 8 else np.float32)
        out.flat[:] = x.flat
        return out
    return np.float64(x) if int(kind) == 8 else F(x)


def fint(x):
    """int(): truncation; a non-finite or out-of-range operand gives what x86's cvttss2si gives a compiled build, the
    "integer indefinite" -2**31 (output.F90:497 relies on such a bin index simply failing its range test)"""
    if isinstance(x, (int, np.integer)):
        return int(x)
    if not np.isfinite(x) or abs(float(x)) >= 2.0 ** 31:
        return -2 ** 31
    return int(x)


INTRINSICS = {"exp": "fexp", "log": "flog", "alog": "flog", "log10": "flog10", "alog10": "flog10", "atan": "fatan", "tan": "ftan",
              "dmod": "fdmod", "dble": "np.float64", "ceiling": "fceiling", "present": "fpresent",
              "aint": "faint", "int": "fint", "real": "freal", "min": "fmin", "max": "fmax", "abs": "abs", "sqrt": "fsqrt", "sum": "fsum",
              "cshift": "fcshift", "mod": "fmod", "modulo": "fmodulo", "float": "F", "nint": "fnint", "floor": "ffloor", "cos": "fcos",
              "sin": "fsin", "sign": "fsign"}
RUNTIME = {"fint": fint, "freal": freal, "fexp": elementary("exp"), "flog": elementary("log"), "flog10": elementary("log10"),
           "fatan": elementary("atan"), "ftan": elementary("tan"), "fdmod": lambda a, b: np.float64(np.fmod(np.float64(a), np.float64(b))),
           "fceiling": lambda x: int(math.ceil(x)), "fpresent": lambda x: x is not None, "F": F, "FArr": FArr, "fdiv": fdiv, "fpow": fpow, "fsum": fsum, "fcshift": fcshift, "frange": frange, "fexit": fexit, "faint": faint,
           "fmodulo": fmodulo, "fmod": fmod, "fmin": fmin, "fmax": fmax, "fsqrt": fsqrt,
           "fsign": lambda a, b: (abs(a) if not np.signbit(b) else -abs(a)) if isinstance(a, np.float64) else (F(abs(a)) if not np.signbit(b) else F(-abs(a))), "fcos": elementary("cos"), "fsin": elementary("sin"), "fnint": lambda x: int(math.copysign(math.floor(abs(float(x)) + 0.5), float(x))), "ffloor": lambda x: int(np.floor(x)), "np": np}
PYKW = {"in", "is", "lambda", "not", "and", "or", "if", "else", "for", "while", "def", "class", "pass", "del", "from", "as", "with"}


# ------------------------------------------------------------------------------------------------------------------
# source handling: cpp conditionals, continuation lines, subroutine extraction
# ------------------------------------------------------------------------------------------------------------------
def preprocess(text, defines):
    out, stack = [], []
    for line in text.splitlines():
        s = line.strip()
        if s.startswith("#"):
            m = re.match(r"#\s*(ifdef|ifndef|else|endif|if|elif|define|undef|include)\b\s*(.*)", s)
            if not m:
                continue
            d, arg = m.group(1), m.group(2).strip()
            if d == "ifdef":
                stack.append(arg.split()[0] in defines)
            elif d == "ifndef":
                stack.append(arg.split()[0] not in defines)
            elif d == "if":
                raise NotImplementedError("#if " + arg)
            elif d == "else":
                stack[-1] = not stack[-1]
            elif d == "endif":
                stack.pop()
            continue
        if all(stack):
            out.append(line)
    return "\n".join(out)


def strip_comment(line):
    q = None
    for i, ch in enumerate(line):
        if q:
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
        elif ch == "!":
            return line[:i]
    return line


def statements(text):
    """comment-free, continuation-joined, lower-cased statements"""
    out, cur = [], ""
    for line in text.splitlines():
        line = strip_comment(line).rstrip()
        if not line.strip():
            continue
        s = line.strip()
        if s.startswith("&"):
            s = s[1:]
        if s.endswith("&"):
            cur += s[:-1] + " "
            continue
        cur += s
        for part in cur.split(";") if ("'" not in cur and '"' not in cur) else [cur]:
            if part.strip():
                st = part.strip().lower()
                # kind-suffixed literals: 1.e7_sprec is a default real, 1._dprec a double
                st = re.sub(r"(?<=[\d.])_sprec\b", "", st)
                st = re.sub(r"(\d+\.?\d*|\.\d+)(?:e([+-]?\d+))?_dprec\b", lambda m: f"{m.group(1)}d{m.group(2) or 0}", st)
                out.append(st)
        cur = ""
    return out


def module_kind(text, name):
    """kind of a module variable, from its declaration in the specification part of the (preprocessed) module text:
    'int' | 'int8' | 'real' | 'real8' | 'logical'"""
    head = text.lower().split("\ncontains")[0]
    probe = Sub.__new__(Sub)
    probe.local, probe.alias, probe.optional = {}, set(), set()
    for st in statements(head):
        try:
            probe.declare(st)
        except SyntaxError:
            continue
    if name.lower() not in probe.local:
        raise KeyError(name)
    return "int8" if name.lower() in getattr(probe, "int8", set()) else probe.local[name.lower()][0]


def extract_subroutine(text, name):
    m = re.search(r"^[ \t]*(?:(?:integer|real|logical)(?:\([a-z0-9_]*\))?\s+)?(subroutine|function)\s+" + name + r"\s*(\(|$)", text, flags=re.I | re.M)
    if not m:
        raise KeyError(name)
    e = re.search(r"^[ \t]*end\s*" + m.group(1) + r"\s+" + name + r"\b", text[m.start():], flags=re.I | re.M)
    e0 = re.search(r"^[ \t]*end\s*" + m.group(1) + r"\b", text[m.start():], flags=re.I | re.M)      # `end subroutine` with no name
    if e is None or e0.end() < e.start():
        e = e0
    return text[m.start():m.start() + e.end()]


# ------------------------------------------------------------------------------------------------------------------
# expression translation (recursive descent over the Fortran expression grammar)
# ------------------------------------------------------------------------------------------------------------------
TOK = re.compile(r"\s*('[^']*'|\"[^\"]*\"|\d+\.(?![a-z]+\.)\d*(?:[ed][+-]?\d+)?|\.\d+(?:[ed][+-]?\d+)?|\d+[ed][+-]?\d+|\d+|\.[a-z]+\.|[a-z_]\w*|\*\*|==|/=|<=|>=|[-+*/(),:<>%=])")


def tokenize(s):
    toks, pos = [], 0
    s = s.strip()
    while pos < len(s):
        m = TOK.match(s, pos)
        if not m:
            raise SyntaxError(f"cannot tokenize {s[pos:]!r} in {s!r}")
        toks.append(m.group(1))
        pos = m.end()
    return toks


class Expr:
    def __init__(self, toks, ctx):
        self.t, self.i, self.ctx = toks, 0, ctx

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else None

    def next(self):
        tok = self.t[self.i]
        self.i += 1
        return tok

    def expect(self, tok):
        if self.next() != tok:
            raise SyntaxError(f"expected {tok} in {' '.join(self.t)}")

    def parse(self):
        r = self.p_or()
        if self.peek() is not None:
            raise SyntaxError(f"trailing {self.peek()!r} in {' '.join(self.t)}")
        return r

    def p_or(self):
        l = self.p_and()
        while self.peek() == ".or.":
            self.next(); l = f"({l} or {self.p_and()})"
        return l

    def p_and(self):
        l = self.p_not()
        while self.peek() == ".and.":
            self.next(); l = f"({l} and {self.p_not()})"
        return l

    def p_not(self):
        if self.peek() == ".not.":
            self.next()
            return f"(not {self.p_not()})"
        return self.p_rel()

    REL = {".le.": "<=", ".lt.": "<", ".ge.": ">=", ".gt.": ">", ".eq.": "==", ".ne.": "!=", "==": "==", "/=": "!=", "<=": "<=", ">=": ">=",
           "<": "<", ">": ">"}

    def p_rel(self):
        l = self.p_add()
        if self.peek() in self.REL:
            op = self.REL[self.next()]
            l = f"({l} {op} {self.p_add()})"
        return l

    def p_add(self):
        if self.peek() in ("-", "+"):
            op = self.next()
            l = f"({op}{self.p_mul()})"
        else:
            l = self.p_mul()
        while self.peek() in ("+", "-"):
            op = self.next()
            l = f"({l} {op} {self.p_mul()})"
        return l

    def p_mul(self):
        l = self.p_pow()
        while self.peek() in ("*", "/"):
            op = self.next()
            r = self.p_pow()
            l = f"({l} * {r})" if op == "*" else f"fdiv({l}, {r})"
        return l

    def p_pow(self):
        b = self.p_unary()
        if self.peek() == "**":
            self.next()
            return f"fpow({b}, {self.p_pow()})"
        return b

    def p_unary(self):
        if self.peek() in ("-", "+"):
            op = self.next()
            return f"({op}{self.p_unary()})"
        return self.p_primary()

    def p_primary(self):
        tok = self.next()
        if tok == "(":
            e = self.p_or()
            self.expect(")")
            r = f"({e})"
        elif tok[0] in "'\"":
            r = repr(tok[1:-1])
        elif tok == ".true.":
            r = "True"
        elif tok == ".false.":
            r = "False"
        elif re.match(r"\d|\.\d", tok):
            if re.fullmatch(r"\d+", tok):
                r = tok
            elif "d" in tok:
                r = f"np.float64({tok.replace('d', 'e')})"
            else:
                r = f"F({tok})"
        elif re.match(r"[a-z_]", tok):
            r = self.p_name(tok)
        else:
            raise SyntaxError(f"unexpected {tok!r} in {' '.join(self.t)}")
        while self.peek() == "%":
            self.next()
            r = f"{r}.{self.next()}"
        return r

    def p_args(self):
        args = []
        if self.peek() == ")":
            self.next()
            return args
        while True:
            # subscript triplet?
            lo = hi = None
            if self.peek() == ":":
                self.next()
                if self.peek() not in (",", ")"):
                    hi = self.p_or()
                args.append(f"slice(None, {hi})")
            else:
                lo = self.p_or()
                if self.peek() == ":":
                    self.next()
                    if self.peek() not in (",", ")", ":"):
                        hi = self.p_or()
                    if self.peek() == ":":                             # lo:hi:stride
                        self.next()
                        args.append(f"slice({lo}, {hi}, {self.p_or()})")
                    else:
                        args.append(f"slice({lo}, {hi})")
                elif self.peek() == "=":                               # keyword argument (cshift(a, shift=1, dim=2))
                    self.next()
                    args.append(f"{lo.split('.')[-1]}={self.p_or()}")
                else:
                    args.append(lo)
            tok = self.next()
            if tok == ")":
                return args
            if tok != ",":
                raise SyntaxError(f"expected , or ) in {' '.join(self.t)}")

    def p_name(self, name):
        ref = self.ctx.ref(name)
        if self.peek() != "(":
            return ref
        self.next()
        args = self.p_args()
        if self.ctx.is_array(name):
            return f"{ref}[{', '.join(args)}]"
        if name in INTRINSICS:
            return f"{INTRINSICS[name]}({', '.join(args)})"
        return f"{ref}({', '.join(args)})"


# ------------------------------------------------------------------------------------------------------------------
# statement translation
# ------------------------------------------------------------------------------------------------------------------
class Sub:
    def __init__(self, source, name, defines=(), global_arrays=(), global_ints=(), alias_globals=(), global_kinds=None):
        """alias_globals: dummy arguments that every caller binds to the module variable of the same name (`dseed`): they are
        read and written as that module variable, which gives functions called inside expressions -- random(dseed) -- the
        by-reference update a plain Python argument cannot have."""
        self.name = name.lower()
        self.local = {}                  # name -> ("int" | "real" | "real8" | "logical", dims or None)
        self.args = []
        self.alias = {a.lower() for a in alias_globals}
        self.global_kinds = dict(global_kinds or {})      # module scalars that are not default real: name -> kind (for READ)
        self.optional = set()
        self.data_inits = []
        self.global_arrays = {a.lower() for a in global_arrays}
        self.global_ints = {a.lower() for a in global_ints}
        text = preprocess(source, set(defines))
        self.stmts = self.forward_gotos(statements(extract_subroutine(text, name)))
        self.py = self.translate()

    # --- names -------------------------------------------------------------------------------------------------
    def pyname(self, n):
        return n + "_" if n in PYKW else n

    def ref(self, n):
        if n in self.local or n in self.args:
            return self.pyname(n)
        if n in INTRINSICS:
            return INTRINSICS[n]
        return "_g." + self.pyname(n)

    def is_array(self, n):
        if n in self.local:
            return self.local[n][1] is not None
        return n in self.global_arrays

    def kind(self, n):
        if n in self.local:
            return self.local[n][0]
        if n in self.global_ints:
            return "int"
        return None

    def ex(self, s):
        return Expr(tokenize(s), self).parse()

    # --- declarations ----------------------------------------------------------------------------------------------
    DECL = re.compile(r"^(integer|real|logical|double precision)\b\s*(\([^)]*\))?\s*((?:,\s*[a-z]+(?:\([^)]*\))?\s*)*)(::)?\s*(.*)$")

    def declare(self, st):
        m = self.DECL.match(st)
        if not m:
            return False
        base, attrs, ents = m.group(1), m.group(3) or "", m.group(5)
        kind = {"integer": "int", "real": "real", "logical": "logical", "double precision": "real8"}[base]
        if base == "real" and re.sub(r"\s|kind=", "", m.group(2) or "") in ("(dprec)", "(8)"):
            kind = "real8"
        wide_int = base == "integer" and re.sub(r"\s|kind=", "", m.group(2) or "") == "(8)"
        dims = None
        dm = re.search(r"dimension\s*\(([^)]*(?:\([^)]*\)[^)]*)*)\)", attrs)
        if dm:
            dims = dm.group(1)
        # split the entity list on top-level commas
        parts, depth, cur = [], 0, ""
        for ch in ents:
            if ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
            if ch == "," and depth == 0:
                parts.append(cur); cur = ""
            else:
                cur += ch
        if cur.strip():
            parts.append(cur)
        for p_ in parts:
            p_ = p_.strip().split("=")[0].strip()
            em = re.match(r"([a-z_]\w*)\s*(?:\((.*)\))?$", p_)
            if not em:
                raise SyntaxError("declaration entity " + p_)
            if em.group(1) in self.alias:
                continue
            self.local[em.group(1)] = (kind, em.group(2) or dims)
            if wide_int:
                self.int8 = getattr(self, "int8", set()) | {em.group(1)}
            if "optional" in attrs:
                self.optional.add(em.group(1))
        return True

    # --- statements ------------------------------------------------------------------------------------------------
    def assign(self, lhs, rhs):
        lhs = lhs.strip()
        r = self.ex(rhs)
        m = re.match(r"([a-z_]\w*)\s*$", lhs)
        if m:
            n = m.group(1)
            if self.is_array(n):
                return f"{self.ref(n)}.set({r})"
            k = self.kind(n)
            if k == "int":
                return f"{self.ref(n)} = fint({r})"
            if k == "real":
                return f"{self.ref(n)} = F({r})"
            if k == "real8":
                return f"{self.ref(n)} = np.float64({r})"
            if k == "logical":
                return f"{self.ref(n)} = bool({r})"
            return f"{self.ref(n)} = fassign_global({r})"
        return f"{self.ex(lhs)} = {r}"

    def split_assign(self, st):
        depth = 0
        for i, ch in enumerate(st):
            if ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
            elif ch == "=" and depth == 0 and st[i - 1] not in "<>/=" and st[i + 1:i + 2] != "=":
                return st[:i], st[i + 1:]
        return None

    def simple(self, st):
        if st.startswith("call mpi_sendrecv"):
            # one rank on a periodic axis exchanges with itself (plusrank == minusrank == rank, e.g. fieldboundaries.F90:
            # 1168-1181): MPI_SendRecv(sendbuf, ..., recvbuf, ...) is then "recvbuf = sendbuf"
            # With several ranks (Globals.comm set) the message goes through Comm; either way the transfer is `count` elements
            # in Fortran order (Payload).
            a = self.split_dims(st[st.index("(") + 1:st.rindex(")")])
            return self.assign(a[5], f"mpi_xchg({a[0]}, {a[1]}, {a[3]}, {a[4]}, {a[8]}, {a[9]}, {a[2]}, {a[7]})")
        if st.startswith("call mpi_allgather"):
            a = self.split_dims(st[st.index("(") + 1:st.rindex(")")])
            return self.assign(a[3], f"mpi_allgather({a[0]})")
        if st.startswith("call mpi_allreduce"):
            a = self.split_dims(st[st.index("(") + 1:st.rindex(")")])
            return self.assign(a[1], f"mpi_allreduce({a[0]}, {a[2]}, {a[4]})")
        if st.startswith("call mpi_type_create_subarray"):
            # the derived datatypes of filter2's ghost exchange (fields.F90:1449-1600): keep the box the reference describes
            a = self.split_dims(st[st.index("(") + 1:st.rindex(")")])
            return self.assign(a[6], f"mpi_subarray({a[1]}, {a[2]}, {a[3]})")
        m = re.match(r"allocate\s*\((.*)\)$", st)
        if m:
            outl = []
            for ent in self.split_dims(m.group(1)):
                em = re.match(r"([a-z_]\w*)\s*\((.*)\)$", ent.strip())
                dd = ", ".join(self.ex(d) for d in self.split_dims(em.group(2)))
                k = self.kind(em.group(1))
                outl.append(f"{self.ref(em.group(1))} = FArr(({dd},), {'np.int64' if k == 'int' else 'np.float32'})")
            return "; ".join(outl)
        if st.startswith("deallocate"):
            return "pass"
        if st.startswith("call mpi_"):
            return "pass"
        if st.startswith("call "):
            m = re.match(r"call\s+([a-z_]\w*)\s*(?:\((.*)\))?$", st)
            args = Expr(tokenize("(" + (m.group(2) or "") + ")"), self)
            args.next()
            a = args.p_args()
            call = f"_g.{m.group(1)}({', '.join(a)})"
            # scalar actual arguments that are variables get the callee's final value of the dummy (by-reference semantics)
            back = []
            for i, raw in enumerate(self.split_dims(m.group(2) or "")):
                raw = raw.strip()
                if "=" in raw and not re.search(r"[<>=/]=|==", raw):
                    continue                                           # keyword argument
                idm = re.fullmatch(r"[a-z_]\w*", raw)
                if idm and not self.is_array(raw) and raw not in INTRINSICS:
                    k = self.kind(raw)
                    cast = {"int": "int", "real": "F", "real8": "np.float64", "logical": "bool"}.get(k, "")
                    back.append(f"{self.ref(raw)} = {cast}(_r[{i}])")
                elif re.fullmatch(r"[a-z_]\w*\s*\(.*\)\s*%\s*[a-z_]\w*", raw) or \
                        (re.fullmatch(r"([a-z_]\w*)\s*\(([^:]*)\)", raw) and self.is_array(raw.split("(")[0].strip())):
                    back.append(f"{self.ex(raw)} = _r[{i}]")
            if not back:
                return call
            return f"_r = {call}\n" + "".join("@IND@if _r is not None: " + b + "\n" for b in back).rstrip("\n")
        if st in ("return",):
            return f"return {self.result()}"
        if st in ("continue",):
            return "pass"
        if st == "cycle":
            return "continue"
        if st == "exit":
            return "break"
        m = re.match(r"read\s*\(\s*(\d+)\s*\)\s*(.+)$", st)
        if m:                                                          # unformatted sequential READ(unit) io-list
            lines = [f"_rd = _g.fread({m.group(1)})"]

            def target(t, depth):
                t = t.strip()
                pad = "    " * depth
                parts = self.split_dims(t[1:-1]) if t.startswith("(") and t.endswith(")") else []
                if len(parts) >= 3 and re.match(r"[a-z_]\w*\s*=", parts[-2]):          # implied DO, possibly nested
                    var, lo = (v.strip() for v in parts[-2].split("=", 1))
                    lines.append(f"{pad}for {self.ref(var)} in frange({self.ex(lo)}, {self.ex(parts[-1])}):")
                    for inner in parts[:-2]:
                        target(inner, depth + 1)
                    return
                mc = re.fullmatch(r"([a-z_]\w*)\s*\((.*)\)\s*%\s*([a-z_]\w*)", t)
                if mc:
                    lines.append(f"{pad}_rd.comp({self.ref(mc.group(1))}, {self.ex(mc.group(2))}, {mc.group(3)!r})")
                    return
                ma = re.fullmatch(r"([a-z_]\w*)\s*\((.*)\)", t)
                if ma and self.is_array(ma.group(1)):
                    idx = ", ".join(self.ex(e) for e in self.split_dims(ma.group(2)))
                    lines.append(f"{pad}_rd.elem({self.ref(ma.group(1))}, ({idx},))")
                    return
                k = self.kind(t) or self.global_kinds.get(t, "real")
                lines.append(f"{pad}{self.ref(t)} = _rd.scalar({k!r})")
            for it in self.split_dims(m.group(2)):
                target(it, 0)
            lines.append("_rd.close()")
            return "\n".join(lines)
        m = re.match(r"write\s*\(\s*(\d+)\s*\)\s*(.+)$", st)
        if m:                                                          # unformatted sequential WRITE(unit) io-list
            items = []
            for it in self.split_dims(m.group(2)):
                it = it.strip()
                parts = self.split_dims(it[1:-1]) if it.startswith("(") and it.endswith(")") else []
                if len(parts) >= 3 and re.match(r"[a-z_]\w*\s*=", parts[-2]):          # implied DO: (expr..., n=lo,hi)
                    var, lo = (v.strip() for v in parts[-2].split("=", 1))
                    exprs = ", ".join(self.ex(e) for e in parts[:-2])
                    items.append(f"[[{exprs}] for {self.ref(var)} in frange({self.ex(lo)}, {self.ex(parts[-1])})]")
                else:
                    items.append(self.ex(it))
            return f"_g.fwrite({m.group(1)}, [{', '.join(items)}])"
        if st.startswith("print") or st.startswith("write") or st.startswith("stop") or re.match(r"(open|close|rewind)\b", st):
            return "pass"
        m = re.match(r"go\s*to\s+(\d+)$", st)
        if m:
            if m.group(1) not in self.cyc:
                raise SyntaxError("goto that is neither a forward skip nor a cycle: " + st)
            return "continue"
        if re.match(r"\d+\s+continue$", st):
            return "pass"
        sa = self.split_assign(st)
        if sa:
            return self.assign(*sa)
        raise SyntaxError("statement: " + st)

    @staticmethod
    def forward_gotos(stmts):
        """An unconditional forward `goto N` ... `N continue` (fields.F90:1192-1210 disables a block this way): the statements in
        between are never executed, so they are dropped.  Anything else with a goto is refused."""
        out, i = [], 0
        cyc = Sub.cycle_labels(stmts)
        while i < len(stmts):
            m = re.match(r"go\s*to\s+(\d+)$", stmts[i])
            if m and m.group(1) not in cyc:
                lab = re.compile(m.group(1) + r"\s+continue$")
                j = next((k for k in range(i + 1, len(stmts)) if lab.match(stmts[k])), None)
                if j is None:
                    raise SyntaxError("goto without a forward label: " + stmts[i])
                i = j + 1
                continue
            out.append(stmts[i])
            i += 1
        return out

    @staticmethod
    def cycle_labels(stmts):
        """labels of `N continue` statements that are the last statement of a do loop: a `go to N` from inside that loop
        (particles_movedeposit.F90:1652 `if(in) go to 58`) is a CYCLE"""
        return {m.group(1) for a, b in zip(stmts, stmts[1:]) if (m := re.match(r"(\d+)\s+continue$", a)) and b in ("enddo", "end do")}

    def translate(self):
        self.cyc = self.cycle_labels(self.stmts)
        hdr = self.stmts[0]
        m = re.match(r"(?:(integer|real|logical)(?:\([a-z0-9_]*\))?\s+)?(subroutine|function)\s+([a-z_]\w*)\s*(?:\((.*)\))?", hdr)
        self.dummies = [a.strip() for a in (m.group(4) or "").split(",") if a.strip()]
        self.args = [a for a in self.dummies if a not in self.alias]
        self.is_function = m.group(2) == "function"
        if self.is_function:             # the result variable carries the function's name and type
            self.local[self.name] = ({"integer": "int", "real": "real", "logical": "logical"}[m.group(1) or "real"], None)
        body, ind = [], 1
        decls_done = []
        loops, nloop = [], [0]

        def emit(s):
            for ln in s.split("\n"):
                body.append("    " * ind + ln.replace("@IND@", ""))
        for st in self.stmts[1:]:
            if re.match(r"end\s*(subroutine|function)\b", st):
                break
            if st.startswith("implicit") or st.startswith("use ") or st.startswith("intent") or st.startswith("external") \
                    or st.startswith("character"):
                continue
            if self.declare(st):
                continue
            if st.startswith("data "):                               # DATA a/1.d0/, b/2.d0/
                self.data_inits += re.findall(r"([a-z_]\w*)\s*/\s*([^/]+?)\s*/", st[5:])
                continue
            # allocate local arrays lazily, once, at the first executable statement
            if not decls_done:
                decls_done.append(1)
                for n, (k, dims) in self.local.items():
                    if n in self.args:
                        continue
                    if dims is not None and ":" in dims and not re.search(r"\w\s*:|:\s*\w", dims):
                        emit(f"{self.pyname(n)} = None")                      # allocatable: created by `allocate`
                    elif dims is not None:
                        dd = ", ".join(self.ex(d) for d in self.split_dims(dims))
                        emit(f"{self.pyname(n)} = FArr(({dd},), {'np.int64' if k == 'int' else 'np.float64' if k == 'real8' else 'np.float32'})")
                    elif k == "real":
                        emit(f"{self.pyname(n)} = F(0.0)")
                    elif k == "real8":
                        emit(f"{self.pyname(n)} = np.float64(0.0)")
                    elif k == "int":
                        emit(f"{self.pyname(n)} = 0")
                    else:
                        emit(f"{self.pyname(n)} = False")
                for lhs_, rhs_ in self.data_inits:
                    emit(self.assign(lhs_, rhs_))
            m = re.match(r"if\s*\((.*)\)\s*then$", st)
            if m:
                emit(f"if {self.ex(m.group(1))}:"); ind += 1; continue
            m = re.match(r"else\s*if\s*\((.*)\)\s*then$", st)
            if m:
                ind -= 1; emit(f"elif {self.ex(m.group(1))}:"); ind += 1; continue
            if st == "else":
                ind -= 1; emit("else:"); ind += 1; continue
            if st in ("endif", "end if"):
                emit("pass"); ind -= 1; continue
            m = re.match(r"select\s*case\s*\((.*)\)$", st)
            if m:
                emit(f"_sel = {self.ex(m.group(1))}"); emit("if False:"); ind += 1; continue
            m = re.match(r"case\s*\((.*)\)$", st)
            if m:
                emit("pass"); ind -= 1; emit(f"elif _sel in ({', '.join(self.ex(v) for v in self.split_dims(m.group(1)))},):"); ind += 1; continue
            if st == "case default":
                emit("pass"); ind -= 1; emit("else:"); ind += 1; continue
            if st in ("end select", "endselect"):
                emit("pass"); ind -= 1; continue
            m = re.match(r"where\s*\((.*)\)$", st)
            if m:
                emit(f"_w = {self.ex(m.group(1))}"); self.in_where = True; continue
            if st == "elsewhere":
                emit("_w = ~_w"); continue
            if st in ("endwhere", "end where"):
                self.in_where = False; continue
            if getattr(self, "in_where", False):
                lhs, rhs = self.split_assign(st)
                with_err = f"{self.ref(lhs.strip())}.where_set(_w, {self.ex(rhs)})"
                emit("with np.errstate(all='ignore'):"); ind += 1; emit(with_err); ind -= 1
                continue
            m = re.match(r"do\s+while\s*\((.*)\)$", st)
            if m:
                emit(f"while {self.ex(m.group(1))}:"); ind += 1; loops.append((None, None)); continue
            m = re.match(r"do\s+([a-z_]\w*)\s*=\s*(.*)$", st)
            if m:
                parts = self.split_dims(m.group(2))
                # Fortran evaluates the bounds once and leaves the variable at its first failing value after the loop
                # (optimized_filters.F90 relies on that: "i = ntimes-1" after `do i=2,ntimes-2,2`)
                nloop[0] += 1
                b = f"_b{nloop[0]}"
                emit(f"{b} = ({', '.join(self.ex(p_) for p_ in parts)},)")
                emit(f"for {self.ref(m.group(1))} in frange(*{b}):")
                loops.append((self.ref(m.group(1)), b))
                ind += 1; continue
            if st in ("enddo", "end do"):
                emit("pass"); ind -= 1
                var, b = loops.pop()
                if var is not None:
                    emit(f"{var} = fexit(*{b})")
                continue
            m = re.match(r"if\s*\(", st)
            if m:
                # one-line if: find the matching parenthesis
                depth, j = 0, st.index("(")
                for j in range(st.index("("), len(st)):
                    depth += st[j] == "("
                    depth -= st[j] == ")"
                    if depth == 0:
                        break
                cond, rest = st[st.index("(") + 1:j], st[j + 1:].strip()
                emit(f"if {self.ex(cond)}:"); ind += 1; emit(self.simple(rest)); ind -= 1
                continue
            emit(self.simple(st))
        args = ", ".join(["_g"] + [("_alias_" + a if a in self.alias else self.pyname(a)) + ("=None" if a in self.optional else "")
                                   for a in self.dummies])
        body.append(f"    return {self.result()}")
        return f"def {self.name}({args}):\n" + "\n".join(body or ["    pass"]) + "\n"

    def result(self):
        """a function returns its result variable; a subroutine returns its dummies in order, so that the caller can copy
        scalar results back into its actual arguments (Fortran passes by reference)"""
        if self.is_function:
            return self.pyname(self.name)
        return "(" + "".join(self.ref(a) + ", " for a in self.dummies) + ")"

    @staticmethod
    def split_dims(s):
        parts, depth, cur = [], 0, ""
        for ch in s:
            if ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
            if ch == "," and depth == 0:
                parts.append(cur.strip()); cur = ""
            else:
                cur += ch
        if cur.strip():
            parts.append(cur.strip())
        return parts

    def compile(self):
        ns = dict(RUNTIME)
        ns["fassign_global"] = lambda v: v
        exec(compile(self.py, f"<f90:{self.name}>", "exec"), ns)
        return ns[self.name]


class Globals:
    """module variables of the reference (m_fields, m_particles, ...) as attributes"""

    comm = None
    sprec, dprec = 4, 8                  # the kind parameters of the reference (real(x, sprec), real(x, dprec))

    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)

    def mpi_xchg(self, sendbuf, count, dest, sendtag, source, recvtag, sendtype=None, recvtype=None):
        pay = as_payload(sendbuf, count, sendtype)
        if self.comm is not None:         # one rank: every neighbour is the rank itself
            pay = self.comm.sendrecv(int(self.rank), pay, dest, sendtag, source, recvtag)
        return Payload(pay.data, recvtype if isinstance(recvtype, Subarray) else None)

    def fread(self, unit):
        """the next unformatted record of `unit` (self.units[unit]: list of marker-framed records) as a reader"""
        if not hasattr(self, "_rpos"):
            self._rpos = {}
        i = self._rpos.get(int(unit), 0)
        self._rpos[int(unit)] = i + 1
        if not hasattr(self, "readers"):
            self.readers = {}
        self.readers[int(unit)] = RecordReader(self.units[int(unit)][i])     # kept: .left = bytes the READ did not consume
        return self.readers[int(unit)]

    def fwrite(self, unit, items):
        """one unformatted sequential record as gfortran lays it out: 4-byte length, the items back to back in their own
        kinds (default integer and real 4 bytes, double precision 8, arrays in storage order), 4-byte length"""
        def enc(v):
            if isinstance(v, (list, tuple)):
                return b"".join(enc(x) for x in v)
            if isinstance(v, FArr):
                return np.ascontiguousarray(v.flat).tobytes()
            if isinstance(v, (bool, np.bool_)):
                return np.int32(bool(v)).tobytes()
            if isinstance(v, (int, np.integer)):
                return np.int32(v).tobytes()
            if isinstance(v, np.float64):
                return v.tobytes()
            if isinstance(v, (np.float32, float)):
                return np.float32(v).tobytes()
            raise TypeError(f"cannot write {type(v)}")
        body = enc(items)
        mark = np.int32(len(body)).tobytes()
        if not hasattr(self, "units"):
            self.units = {}
        self.units.setdefault(int(unit), []).append(mark + body + mark)

    def mpi_allgather(self, val):
        """MPI_Allgather of one value per rank into an array indexed by rank"""
        vals = [val] if self.comm is None else self.comm.allreduce(int(self.rank), val, "gather")
        return Payload(np.array(vals))

    def mpi_allreduce(self, sendbuf, count, op):
        """MPI_Allreduce(sendbuf, recvbuf, count, type, op, ...): every contribution is logged as it crosses the MPI boundary
        (allreduce_log) -- the per-rank value the reference computed; ranks are combined in rank order"""
        val = as_payload(sendbuf, count).data if isinstance(sendbuf, (FArr, np.ndarray)) else sendbuf
        if not hasattr(self, "allreduce_log"):
            self.allreduce_log = []
        self.allreduce_log.append(np.array(val).copy())
        if self.comm is not None:
            val = self.comm.allreduce(int(self.rank), val, op)
        return Payload(np.atleast_1d(val)) if isinstance(sendbuf, (FArr, np.ndarray)) else val

    @staticmethod
    def mpi_subarray(sizes, subsizes, starts):
        return Subarray(sizes.flat, subsizes.flat, starts.flat)
