"""Build + bind the generated CUDA libraries (ctypes over ``include/cfem.h``).

The reference reaches its evaluation code through ``compile_class()`` / a
written-and-imported generated module (/root/reference/attas_sp_ml.py:85-86,
/root/reference/mc_blackbox_cfem.py:81-95).  Here the generated artefact is a
shared library: ``codegen`` emits one ``.cu`` per problem structure, ``nvcc``
compiles it for sm_100a into ``colloc_fem_code_b200/_gen/`` (in-tree, cached by
a hash of the generated source and the hand-written headers, the same
write-once/reuse pattern as the reference's ``Generated*.py`` files) and
``ctypes`` binds the C ABI.

There is deliberately no fallback: if the library cannot be built or loaded, or
no CUDA device is present, evaluation raises :class:`CfemError`.
"""

import ctypes
import hashlib
import json
import os
import shutil
import subprocess
import threading

import numpy as np

from . import codegen

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
INCLUDE = os.path.join(os.path.dirname(HERE), 'include')
GEN_DIR = os.environ.get('CFEM_GEN_DIR') or os.path.join(HERE, '_gen')

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo',
              '-O3', '-std=c++17', '-Xcompiler', '-fPIC']
LINK_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-Xcompiler',
              '-fPIC']

F, GRAD, G, JAC, HESS, ALL = 1, 2, 4, 8, 16, 31
X, LAMBDA = 32, 64

_c_double_p = ctypes.POINTER(ctypes.c_double)
_c_int64_p = ctypes.POINTER(ctypes.c_int64)

#: every symbol include/cfem.h declares: name -> (restype, argtypes)
ABI = {
    'cfem_abi_version': (ctypes.c_int, []),
    'cfem_model_json': (ctypes.c_char_p, []),
    'cfem_model_sizes': (ctypes.c_int, [ctypes.c_int64, ctypes.c_int32,
                                        _c_int64_p, _c_int64_p, _c_int64_p,
                                        _c_int64_p]),
    'cfem_create': (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p),
                                   ctypes.c_int64, ctypes.c_int32,
                                   ctypes.c_int32,
                                   ctypes.POINTER(ctypes.c_void_p),
                                   ctypes.c_int32, ctypes.c_void_p,
                                   ctypes.c_int32, ctypes.c_int32]),
    'cfem_destroy': (None, [ctypes.c_void_p]),
    'cfem_last_error': (ctypes.c_char_p, [ctypes.c_void_p]),
    'cfem_set_stream': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    'cfem_sizes': (ctypes.c_int, [ctypes.c_void_p, _c_int64_p, _c_int64_p,
                                  _c_int64_p, _c_int64_p]),
    'cfem_layout': (ctypes.c_int, [ctypes.c_void_p] + [ctypes.c_void_p] * 5),
    'cfem_set_dvec': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    'cfem_set_dvec_device': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    'cfem_set_multipliers': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_double,
                                            ctypes.c_void_p]),
    'cfem_set_multipliers_device': (ctypes.c_int, [ctypes.c_void_p,
                                                   ctypes.c_double,
                                                   ctypes.c_void_p]),
    'cfem_eval': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint32]),
    'cfem_fetch': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint32,
                                  ctypes.c_void_p]),
    'cfem_fetch_async': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint32,
                                        ctypes.c_void_p]),
    'cfem_io_layout': (ctypes.c_int, [ctypes.c_void_p, _c_int64_p, _c_int64_p,
                                      _c_int64_p, _c_int64_p]),
    'cfem_set_inputs': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint32,
                                       ctypes.c_double, ctypes.c_void_p]),
    'cfem_fetch_results_async': (ctypes.c_int, [ctypes.c_void_p,
                                                ctypes.c_uint32,
                                                ctypes.c_void_p]),
    'cfem_eval_callback_set': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_double,
                                              ctypes.c_void_p,
                                              ctypes.c_void_p]),
    'cfem_upload_pieces': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint32,
                                          ctypes.c_void_p, ctypes.c_int32,
                                          ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_void_p]),
    'cfem_fetch_pieces': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint32,
                                         ctypes.c_void_p, ctypes.c_int32,
                                         ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_void_p]),
    'cfem_set_obj_factor': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_double]),
    'cfem_eval_f': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    'cfem_eval_grad_f': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    'cfem_eval_g': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    'cfem_eval_jac_values': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    'cfem_eval_hess_values': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_double,
                                             ctypes.c_void_p,
                                             ctypes.c_void_p]),
    'cfem_device_ptrs': (ctypes.c_int, [ctypes.c_void_p]
                         + [ctypes.POINTER(ctypes.c_void_p)] * 8),
    'cfem_apply_reduced': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    'cfem_set_peers': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32,
                                      ctypes.c_int32,
                                      ctypes.POINTER(ctypes.c_void_p),
                                      ctypes.POINTER(ctypes.c_void_p)]),
    'cfem_peer_layout': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32,
                                        _c_int64_p, _c_int64_p]),
    'cfem_set_peer_mode': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32]),
    'cfem_synchronize': (ctypes.c_int, [ctypes.c_void_p]),
    'cfem_set_graph_mode': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32]),
    'cfem_event_record': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32]),
    'cfem_event_elapsed_ms': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32,
                                             ctypes.c_int32,
                                             ctypes.POINTER(ctypes.c_float)]),
    'cfem_set_kernel_timing': (ctypes.c_int, [ctypes.c_void_p,
                                              ctypes.c_int32]),
    'cfem_last_sample_kernel_ms': (ctypes.c_int,
                                   [ctypes.c_void_p,
                                    ctypes.POINTER(ctypes.c_float)]),
    'cfem_sample_kernel_ms_history': (ctypes.c_int,
                                      [ctypes.c_void_p,
                                       ctypes.POINTER(ctypes.c_float),
                                       ctypes.c_int32]),
    'cfem_launch_count': (ctypes.c_int64, [ctypes.c_void_p]),
    'cfem_flush_l2': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t]),
    'cfem_host_alloc': (ctypes.c_void_p, [ctypes.c_size_t]),
    'cfem_host_free': (None, [ctypes.c_void_p]),
    'cfem_host_register': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t]),
    'cfem_host_unregister': (ctypes.c_int, [ctypes.c_void_p]),
    'cfem_store_release_i64': (None, [ctypes.c_void_p, ctypes.c_int64]),
    'cfem_load_acquire_i64': (ctypes.c_int64, [ctypes.c_void_p]),
}


class CfemError(RuntimeError):
    """A C-ABI call failed (or the CUDA library is unavailable)."""


# ----------------------------------------------------------------------------
# building
# ----------------------------------------------------------------------------

def _read(path):
    with open(path, 'rb') as fh:
        return fh.read()


def _digest(*parts):
    h = hashlib.sha256()
    for part in parts:
        h.update(part if isinstance(part, bytes) else part.encode())
    return h.hexdigest()[:16]


_build_lock = threading.Lock()


def nvcc_path():
    path = os.environ.get('CFEM_NVCC') or shutil.which('nvcc') \
        or '/usr/local/cuda/bin/nvcc'
    if not os.path.isfile(path):
        raise CfemError('nvcc not found: cannot build the CUDA model library')
    return path


def _nvcc(args, what):
    proc = subprocess.run([nvcc_path()] + args, capture_output=True, text=True)
    if proc.returncode != 0:
        raise CfemError(f'nvcc failed for {what}:\n' + proc.stdout
                        + proc.stderr)
    return proc.stderr


def build_library(structure, label='model', verbose=False, files=None,
                  **gen_kwargs):
    """Generate + compile the library of ``structure``; returns the .so path.

    The translation unit(s) of ``codegen.Generator.sources`` are compiled to
    objects and linked; objects and the library are cached in ``GEN_DIR`` by a
    hash of their source, the hand-written files they include and the flags.
    """
    src = codegen.generate(structure, **gen_kwargs)
    flags = ' '.join(NVCC_FLAGS)
    hand = [_read(os.path.join(CSRC, n)) for n in
            ('cfem_args.cuh', 'cfem_device.cuh', 'cfem_host.inl')] \
        + [_read(os.path.join(INCLUDE, 'cfem.h'))]
    units = sorted(src)
    keys = {u: _digest(src[u], flags, *hand) for u in units}
    so_path = os.path.join(
        GEN_DIR, f'cfem_{label}_' + ''.join(keys[u] for u in units) + '.so')
    if files is not None:       # every artefact this library is made of
        files.append(so_path)
        for unit in units:
            stem = os.path.join(GEN_DIR, f'cfem_{label}_{unit}_{keys[unit]}')
            files += [stem + '.o', stem + '.cu']
    if os.path.isfile(so_path):
        return so_path
    with _build_lock:
        if os.path.isfile(so_path):
            return so_path
        os.makedirs(GEN_DIR, exist_ok=True)
        objs = []
        for unit in units:
            stem = os.path.join(GEN_DIR, f'cfem_{label}_{unit}_{keys[unit]}')
            objs.append(stem + '.o')
            if os.path.isfile(stem + '.o'):
                continue
            with open(stem + '.cu', 'w') as fh:
                fh.write(src[unit])
            tmp = f'{stem}.tmp{os.getpid()}.o'
            log = _nvcc(NVCC_FLAGS + (['-Xptxas=-v'] if verbose else [])
                        + ['-c', '-I', INCLUDE, '-I', CSRC, '-o', tmp,
                           stem + '.cu'], stem + '.cu')
            if verbose:
                print(log)
            os.replace(tmp, stem + '.o')
        tmp = f'{so_path}.tmp{os.getpid()}'
        _nvcc(LINK_FLAGS + ['-shared', '-o', tmp] + objs, so_path)
        os.replace(tmp, so_path)
    return so_path


def structure_label(structure):
    dims = '_'.join(f"{v['name']}{v['core']}" for v in structure.vars
                    if v['per_sample'])
    names = [f['name'] for f in structure.funs]
    h = hashlib.sha256(','.join(names).encode()).hexdigest()[:6]
    return f'{dims}_{len(names)}f{h}'


class Library:
    """A loaded model library with typed entry points."""

    _loaded = {}

    def __init__(self, path):
        self.path = path
        try:
            self.dll = ctypes.CDLL(path)
        except OSError as exc:
            raise CfemError(f'cannot load CUDA model library {path}: {exc}')
        for name, (restype, argtypes) in ABI.items():
            try:
                fn = getattr(self.dll, name)
            except AttributeError:
                raise CfemError(f'{path} does not export {name}')
            fn.restype = restype
            fn.argtypes = argtypes
            setattr(self, name, fn)
        if self.cfem_abi_version() != 1:
            raise CfemError(f'{path}: unsupported ABI version')
        self.model = json.loads(self.cfem_model_json().decode())

    @classmethod
    def load(cls, path):
        if path not in cls._loaded:
            cls._loaded[path] = cls(path)
        return cls._loaded[path]

    @classmethod
    def for_structure(cls, structure, **kwargs):
        return cls.load(build_library(structure, structure_label(structure),
                                      **kwargs))


# ----------------------------------------------------------------------------
# one problem (or batch of same-shaped problems) on one GPU
# ----------------------------------------------------------------------------

def _as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Handle:
    """Thin object wrapper of a ``cfem_problem*``."""

    def __init__(self, lib, n_samples, data, scalars, batch=1, halo=0,
                 device=0):
        self.lib = lib
        self.batch = int(batch)
        self._ptr = ctypes.c_void_p()
        self._keep = [_as_f64(d) for d in data]
        arr = (ctypes.c_void_p * max(1, len(data)))(
            *[d.ctypes.data for d in self._keep])
        sc = _as_f64(scalars)
        rc = lib.cfem_create(ctypes.byref(self._ptr), int(n_samples),
                             self.batch, int(halo), arr, len(data),
                             sc.ctypes.data if sc.size else None,
                             sc.size, int(device))
        if rc != 0:
            msg = lib.cfem_last_error(None).decode()
            self._ptr = ctypes.c_void_p()
            raise CfemError(f'cfem_create failed ({rc}): {msg}')
        sizes = [ctypes.c_int64() for _ in range(4)]
        self._check(lib.cfem_sizes(self._ptr, *[ctypes.byref(s)
                                                for s in sizes]))
        self.ndec, self.ncons, self.nnz_jac, self.nnz_hess = \
            (s.value for s in sizes)
        self._keep = None       # data was copied to the device

    def _check(self, rc):
        if rc != 0:
            msg = self.lib.cfem_last_error(self._ptr).decode()
            raise CfemError(f'cfem call failed ({rc}): {msg}')

    def close(self):
        if self._ptr:
            self.lib.cfem_destroy(self._ptr)
            self._ptr = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- layout ---------------------------------------------------------------
    def layout(self):
        m = self.lib.model
        ncons = sum(1 for f in m['funs'] if not f['is_objective'])
        arrs = [np.zeros(max(1, n), dtype=np.int64)
                for n in (len(m['vars']), ncons, len(m['jac_blocks']),
                          len(m['hess_blocks']), len(m['funs']))]
        self._check(self.lib.cfem_layout(self._ptr,
                                         *[a.ctypes.data for a in arrs]))
        keys = ('var_offset', 'cons_offset', 'jac_offset', 'hess_offset',
                'fun_rows')
        lens = (len(m['vars']), ncons, len(m['jac_blocks']),
                len(m['hess_blocks']), len(m['funs']))
        return {k: a[:n] for k, a, n in zip(keys, arrs, lens)}

    # -- inputs ---------------------------------------------------------------
    def _shape(self, n):
        return (n,) if self.batch == 1 else (self.batch, n)

    def set_dvec(self, dvec):
        dvec = _as_f64(dvec)
        if dvec.size != self.batch * self.ndec:
            raise ValueError(f'decision vector has {dvec.size} entries, '
                             f'expected {self.batch * self.ndec}')
        self._check(self.lib.cfem_set_dvec(self._ptr, dvec.ctypes.data))
        # the copy is asynchronous on the handle's stream: pageable memory is
        # staged by the runtime before the call returns, pinned memory is not
        self._last_dvec = dvec

    def set_dvec_device(self, dev_ptr):
        self._check(self.lib.cfem_set_dvec_device(self._ptr, int(dev_ptr)))

    def set_multipliers(self, obj_factor, lam):
        lam = _as_f64(lam)
        if lam.size != self.batch * self.ncons:
            raise ValueError(f'multiplier vector has {lam.size} entries, '
                             f'expected {self.batch * self.ncons}')
        self._check(self.lib.cfem_set_multipliers(
            self._ptr, float(obj_factor), lam.ctypes.data))
        self._last_lam = lam

    def set_multipliers_device(self, obj_factor, dev_ptr):
        self._check(self.lib.cfem_set_multipliers_device(
            self._ptr, float(obj_factor), int(dev_ptr)))

    def set_stream(self, stream_ptr):
        self._check(self.lib.cfem_set_stream(self._ptr, stream_ptr))

    # -- evaluation -------------------------------------------------------------
    def eval(self, what):
        self._check(self.lib.cfem_eval(self._ptr, int(what)))

    def fetch(self, which, out=None):
        n = {F: 1, GRAD: self.ndec, G: self.ncons, JAC: self.nnz_jac,
             HESS: self.nnz_hess}[which]
        if out is None:
            out = np.empty(self._shape(n))
        elif (out.dtype != np.float64 or not out.flags.c_contiguous
              or out.size != self.batch * n):
            raise ValueError('output buffer must be C-contiguous float64 of '
                             f'{self.batch * n} entries')
        self._check(self.lib.cfem_fetch(self._ptr, int(which),
                                        out.ctypes.data))
        return out

    def fetch_async(self, which, out):
        """Enqueue the D2H copy of one result into ``out`` (pinned memory);
        valid after :meth:`synchronize`."""
        self._check(self.lib.cfem_fetch_async(self._ptr, int(which),
                                              out.ctypes.data))

    def io_layout(self):
        """Segment offsets / totals (doubles) of the input and result slabs."""
        ino = (ctypes.c_int64 * 2)()
        reo = (ctypes.c_int64 * 5)()
        it, rt = ctypes.c_int64(), ctypes.c_int64()
        self._check(self.lib.cfem_io_layout(
            self._ptr, ino, ctypes.byref(it), reo, ctypes.byref(rt)))
        return list(ino), it.value, list(reo), rt.value

    def set_inputs(self, which, obj_factor, host_block):
        self._check(self.lib.cfem_set_inputs(self._ptr, int(which),
                                             float(obj_factor),
                                             host_block.ctypes.data))

    def fetch_results_async(self, which, host_block):
        self._check(self.lib.cfem_fetch_results_async(
            self._ptr, int(which), host_block.ctypes.data))

    def eval_callback_set(self, obj_factor, host_inputs, host_results):
        """One full callback set between page-locked host blocks, H2D of the
        multipliers overlapped with the D2H of the first results."""
        self._check(self.lib.cfem_eval_callback_set(
            self._ptr, float(obj_factor), host_inputs.ctypes.data,
            host_results.ctypes.data))

    @staticmethod
    def _piece_arrays(pieces):
        arr = np.ascontiguousarray(pieces, dtype=np.int64).reshape(-1, 3)
        return (len(arr), np.ascontiguousarray(arr[:, 0]),
                np.ascontiguousarray(arr[:, 1]),
                np.ascontiguousarray(arr[:, 2]))

    def upload_pieces(self, which, host_vector, pieces):
        """``pieces``: rows (device offset, host offset, length) in doubles."""
        n, dev, host, ln = self._piece_arrays(pieces)
        self._check(self.lib.cfem_upload_pieces(
            self._ptr, int(which), host_vector.ctypes.data, n,
            dev.ctypes.data, host.ctypes.data, ln.ctypes.data))

    def fetch_pieces(self, which, host_vector, pieces):
        n, dev, host, ln = self._piece_arrays(pieces)
        self._check(self.lib.cfem_fetch_pieces(
            self._ptr, int(which), host_vector.ctypes.data, n,
            dev.ctypes.data, host.ctypes.data, ln.ctypes.data))

    def set_obj_factor(self, sigma):
        self._check(self.lib.cfem_set_obj_factor(self._ptr, float(sigma)))

    def synchronize(self):
        self._check(self.lib.cfem_synchronize(self._ptr))

    def device_ptrs(self):
        names = ('dvec', 'lam', 'f', 'grad', 'g', 'jac', 'hess', 'reduce')
        ptrs = [ctypes.c_void_p() for _ in names]
        self._check(self.lib.cfem_device_ptrs(
            self._ptr, *[ctypes.byref(p) for p in ptrs]))
        return {n: p.value for n, p in zip(names, ptrs)}

    def peer_layout(self, world):
        a, b = ctypes.c_int64(), ctypes.c_int64()
        self._check(self.lib.cfem_peer_layout(self._ptr, int(world),
                                              ctypes.byref(a),
                                              ctypes.byref(b)))
        return a.value, b.value

    def set_peers(self, rank, world, inbox_ptrs, flag_ptrs):
        """Enable the in-kernel reduction over peer memory."""
        n = max(1, len(inbox_ptrs))
        ia = (ctypes.c_void_p * n)(*[int(p) for p in inbox_ptrs])
        fa = (ctypes.c_void_p * n)(*[int(p) for p in flag_ptrs])
        self._check(self.lib.cfem_set_peers(self._ptr, int(rank), int(world),
                                            ia, fa))

    def set_graph_mode(self, enabled=True):
        """One CUDA graph launch per evaluation (``cfem_set_graph_mode``)."""
        self._check(self.lib.cfem_set_graph_mode(self._ptr,
                                                 1 if enabled else 0))

    def set_peer_mode(self, pipelined):
        """``True``: post in the kernel, collect on a side stream beside the
        next launch (``cfem_set_peer_mode``)."""
        self._check(self.lib.cfem_set_peer_mode(self._ptr,
                                                1 if pipelined else 0))

    def apply_reduced(self, dev_ptr):
        self._check(self.lib.cfem_apply_reduced(self._ptr, int(dev_ptr)))

    # -- measurement --------------------------------------------------------------
    def event_record(self, slot):
        self._check(self.lib.cfem_event_record(self._ptr, slot))

    def event_elapsed_ms(self, start, stop):
        ms = ctypes.c_float()
        self._check(self.lib.cfem_event_elapsed_ms(self._ptr, start, stop,
                                                   ctypes.byref(ms)))
        return ms.value

    def set_kernel_timing(self, enabled=True):
        self._check(self.lib.cfem_set_kernel_timing(self._ptr, int(enabled)))

    def last_sample_kernel_ms(self):
        ms = ctypes.c_float()
        self._check(self.lib.cfem_last_sample_kernel_ms(self._ptr,
                                                        ctypes.byref(ms)))
        return ms.value

    def sample_kernel_ms_history(self, n):
        n = int(min(n, 64))
        arr = (ctypes.c_float * n)()
        self._check(self.lib.cfem_sample_kernel_ms_history(self._ptr, arr, n))
        return list(arr)

    @property
    def launch_count(self):
        return self.lib.cfem_launch_count(self._ptr)

    def flush_l2(self, nbytes=256 << 20):
        self._check(self.lib.cfem_flush_l2(self._ptr, nbytes))


class PinnedArray:
    """float64 host array in page-locked memory (``cfem_host_alloc``)."""

    def __init__(self, lib, n):
        self.lib = lib
        self.nbytes = max(1, int(n)) * 8
        self.ptr = lib.cfem_host_alloc(self.nbytes)
        if not self.ptr:
            raise CfemError(f'cfem_host_alloc({self.nbytes}) failed')
        buf = (ctypes.c_double * int(n)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=np.float64, count=int(n))

    def close(self):
        if self.ptr:
            self.array = None
            self.lib.cfem_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class HostBuffers:
    """Page-locked host staging for every input and result of a handle: the
    buffers an NLP solver's callbacks read from / write to.  ONE block for the
    inputs and ONE for the results, laid out like the device slabs
    (``cfem_io_layout``), so that a callback set is one H2D and one D2H copy;
    ``dvec, lam, f, grad, g, jac, hess`` are views into them."""

    _RESULTS = ('f', 'grad', 'g', 'jac', 'hess')

    def __init__(self, handle):
        self.handle = handle
        h = handle
        in_off, in_total, res_off, res_total = h.io_layout()
        self._in = PinnedArray(h.lib, in_total)
        self._res = PinnedArray(h.lib, res_total)
        self.inputs, self.results = self._in.array, self._res.array
        self.inputs[:] = 0.0
        B = h.batch
        self.dvec = self.inputs[in_off[0]:in_off[0] + B * h.ndec]
        self.lam = self.inputs[in_off[1]:in_off[1] + B * h.ncons]
        sizes = (B, B * h.ndec, B * h.ncons, B * h.nnz_jac, B * h.nnz_hess)
        for name, off, n in zip(self._RESULTS, res_off, sizes):
            setattr(self, name, self.results[off:off + n])

    def upload(self, obj_factor=None):
        """H2D of the decision vector (and, with ``obj_factor``, of the
        multipliers) in one copy."""
        if obj_factor is None:
            self.handle.set_inputs(X, 0.0, self.inputs)
        else:
            self.handle.set_inputs(X | LAMBDA, obj_factor, self.inputs)

    def fetch(self, which=ALL):
        """D2H of the selected results in one copy, then synchronise."""
        self.handle.fetch_results_async(which, self.results)
        self.handle.synchronize()

    def fetch_all(self):
        self.fetch(ALL)

    def callback_set(self, obj_factor):
        """``upload`` + ``eval(ALL)`` + ``fetch_all`` with the two directions
        of the bus overlapped (``cfem_eval_callback_set``)."""
        self.handle.eval_callback_set(obj_factor, self.inputs, self.results)

    def close(self):
        self.dvec = self.lam = self.inputs = self.results = None
        for name in self._RESULTS:
            setattr(self, name, None)
        self._in.close()
        self._res.close()


class ProblemBackend:
    """CUDA evaluator behind ``optim.Problem`` (one problem, one GPU)."""

    def __init__(self, problem, device=None):
        self.problem = problem
        st = problem.structure
        self.lib = Library.for_structure(st)
        if device is None:
            device = int(os.environ.get('CFEM_DEVICE',
                                        os.environ.get('LOCAL_RANK', 0)))
        self.handle = Handle(self.lib, st.N, [d['source'] for d in st.data],
                             st.scalar_values, device=device)
        h = self.handle
        if (h.ndec, h.ncons) != (problem.ndec, problem.ncons):
            raise CfemError(
                f'layout mismatch: library ({h.ndec}, {h.ncons}) vs problem '
                f'({problem.ndec}, {problem.ncons})')
        if (h.nnz_jac, h.nnz_hess) != (problem.nnzjac, problem.nnzhess):
            raise CfemError('sparsity mismatch between library and problem')

    # IPOPT-shaped entry points.  ``new_x=False`` skips the upload of dvec.
    def set_dvec(self, dvec):
        self.handle.set_dvec(dvec)

    def eval_f(self, dvec=None):
        if dvec is not None:
            self.handle.set_dvec(dvec)
        self.handle.eval(F)
        return float(self.handle.fetch(F)[0])

    def eval_grad_f(self, dvec=None, out=None):
        if dvec is not None:
            self.handle.set_dvec(dvec)
        self.handle.eval(GRAD)
        return self.handle.fetch(GRAD, out)

    def eval_g(self, dvec=None, out=None):
        if dvec is not None:
            self.handle.set_dvec(dvec)
        self.handle.eval(G)
        return self.handle.fetch(G, out)

    def eval_jac_values(self, dvec=None, out=None):
        if dvec is not None:
            self.handle.set_dvec(dvec)
        self.handle.eval(JAC)
        return self.handle.fetch(JAC, out)

    def eval_hess_values(self, dvec, obj_factor, lam, out=None):
        if dvec is not None:
            self.handle.set_dvec(dvec)
        self.handle.set_multipliers(obj_factor, lam)
        self.handle.eval(HESS)
        return self.handle.fetch(HESS, out)

    def eval_all(self, dvec, obj_factor, lam):
        """One fused pass: (f, grad, g, jac values, hess values)."""
        h = self.handle
        h.set_dvec(dvec)
        h.set_multipliers(obj_factor, lam)
        h.eval(ALL)
        return (float(h.fetch(F)[0]), h.fetch(GRAD), h.fetch(G),
                h.fetch(JAC), h.fetch(HESS))

    def close(self):
        self.handle.close()
