"""ctypes binding of libich_b200.so (the C-ABI declared in include/ich_b200.h).

There is NO fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes
import os
import re
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
PKG_ROOT = os.path.dirname(_HERE)
REPO_ROOT = os.path.dirname(PKG_ROOT)
CSRC = os.path.join(PKG_ROOT, 'csrc')
HEADER = os.path.join(REPO_ROOT, 'include', 'ich_b200.h')
LIB_PATH = os.environ.get('ICH_B200_LIB') or os.path.join(_HERE, 'libich_b200.so')     # ICH_B200_LIB: an alternative build (e.g. the -DICH_TC_DEBUG ablation build of scratch/build_debug_lib.sh)
SOURCES = ['api.cu', 'gemm_generic.cu', 'elementwise.cu', 'loss.cu', 'conv_tc.cu', 'conv_tc_stream.cu', 'conv_cin1_tc.cu', 'aux_ops.cu']

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC']

_CTYPE = {'int': ctypes.c_int, 'unsigned char': ctypes.c_ubyte, 'long long': ctypes.c_longlong, 'float': ctypes.c_float, 'double': ctypes.c_double,
          'unsigned int': ctypes.c_uint}


ARG_NAMES = {}       # entry point -> argument names as declared in the header (profiling: algorithmic bytes per launch)


def parse_header(path=HEADER):
    """Return {name: (restype, [argtypes])} for every prototype in the header."""
    text = open(path).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    protos = {}
    for m in re.finditer(r'(const char\*|int)\s+(ich_\w+)\s*\(([^)]*)\)\s*;', text):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        argtypes, names = [], []
        if args and args != 'void':
            for a in args.split(','):
                a = a.strip()
                names.append(re.search(r'(\w+)$', a).group(1))
                if '*' in a:
                    argtypes.append(ctypes.c_void_p)
                else:
                    typ = re.sub(r'\s+\w+$', '', a).replace('const ', '').strip()
                    argtypes.append(_CTYPE[typ])
        protos[name] = (ctypes.c_char_p if 'char' in ret else ctypes.c_int, argtypes)
        ARG_NAMES[name] = names
    return protos


def build(verbose=False, force=False):
    """Compile csrc/*.cu into libich_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cuh')]
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in deps):
        return LIB_PATH
    objdir = os.path.join(PKG_ROOT, 'build')
    os.makedirs(objdir, exist_ok=True)
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    procs = []
    objs = []
    for s in srcs:
        o = os.path.join(objdir, os.path.basename(s) + '.o')
        objs.append(o)
        if not force and os.path.exists(o) and all(os.path.getmtime(o) >= os.path.getmtime(d) for d in [s] + [d for d in deps if d.endswith('.cuh')]):
            continue
        cmd = [nvcc] + NVCC_FLAGS + ['-I', os.path.join(REPO_ROOT, 'include'), '-c', s, '-o', o]
        if verbose:
            print(' '.join(cmd))
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f'nvcc failed for {s}:\n{out}')
        if verbose and out.strip():
            print(out)
    cmd = [nvcc, '-shared', '-o', LIB_PATH] + objs + ['-lcuda', '-lcudart']
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'link failed:\n{r.stdout}')
    return LIB_PATH


_lib = None
_lock = threading.Lock()


def lib():
    """Load (once) and return the ctypes library with typed prototypes. Raises if the .so is absent."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f'ich_b200: {LIB_PATH} is missing -- run `python -c "import __graft_entry__ as g; g.build()"` '
                               '(there is no CPU / PyTorch fallback for the hot path)')
        l = ctypes.CDLL(LIB_PATH)
        for name, (ret, argtypes) in parse_header().items():
            fn = getattr(l, name)        # AttributeError if the header declares a symbol the library lacks
            fn.restype = ret
            fn.argtypes = argtypes
        _lib = l
    return _lib


# kernels launched per entry point (memsets not counted); everything else launches exactly one
KERNELS_PER_CALL = {'ich_bn_act_bwd': 2, 'ich_bn_act_bwd_drop': 2, 'ich_bn_head_bwd': 2, 'ich_seg_loss_fwd': 2, 'ich_tversky_loss_fwd': 2,
                    'ich_infonce_fwd': 2}
LAUNCHES = {}


def launches():
    """Total number of ich_b200 CUDA kernels launched by this process so far."""
    return sum(LAUNCHES.values())


# bench.py sets PROFILE = [] to collect (entry point, {arg name: value}, start event, end event) for every launch (CUDA events on the
# launching stream); None = off (the normal state: no events, no overhead)
PROFILE = None
TAG = None            # set by ops._Timed: 'fwd' / 'dgrad' / 'wgrad' of the conv launch being made


def call(name, *args):
    l = lib()
    LAUNCHES[name] = LAUNCHES.get(name, 0) + KERNELS_PER_CALL.get(name, 1)
    if PROFILE is not None:
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(l, name)(*args)
        e1.record()
        PROFILE.append((name, dict(zip(ARG_NAMES.get(name, ()), args)), TAG, e0, e1))
    else:
        rc = getattr(l, name)(*args)
    if rc != 0:
        raise RuntimeError(f'{name} failed ({rc}): {l.ich_last_error().decode()}')
