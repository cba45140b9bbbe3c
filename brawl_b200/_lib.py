"""ctypes binding of libbrawl_cuda.so (include/brawl_cuda.h).  Fails loudly when the CUDA
library is missing or no device is usable -- there is no CPU fallback in the product path."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BRAWL_CUDA_LIB") or os.path.join(HERE, "libbrawl_cuda.so")   # override: A/B builds of the same ABI


class BrawlCudaError(RuntimeError):
    pass


_lib = None

_vp, _i, _i64, _u64, _d = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_double
_SIGNATURES = {
    "brawl_cuda_version": [],
    "brawl_cuda_device_count": [_vp],
    "brawl_cuda_philox4x32": [_i, _vp, _vp, _vp],
    "brawl_cuda_create": [_i, _i, _i, _i, _i, _i, _vp, _i, _i, _vp],
    "brawl_cuda_destroy": [_vp],
    "brawl_cuda_set_stream": [_vp, _vp],
    "brawl_cuda_synchronize": [_vp],
    "brawl_cuda_info": [_vp, _vp, _vp, _vp, _vp],
    "brawl_cuda_set_config": [_vp, _i, _i, _vp],
    "brawl_cuda_get_config": [_vp, _i, _i, _vp],
    "brawl_cuda_set_lattice": [_vp, _i, _i, _vp],
    "brawl_cuda_get_lattice": [_vp, _i, _i, _vp],
    "brawl_cuda_copy_replica": [_vp, _i, _i],
    "brawl_cuda_random_config": [_vp, _i, _i, _vp, _u64, _u64],
    "brawl_cuda_store_state": [_vp, _i, _i],
    "brawl_cuda_get_order": [_vp, _i, _vp, _i],
    "brawl_cuda_total_energy": [_vp, _i, _i, _i, _vp],
    "brawl_cuda_site_energies": [_vp, _i, _vp],
    "brawl_cuda_nbr_energy": [_vp, _i, _i, _i, _i, _i, _vp],
    "brawl_cuda_pair_dE": [_vp, _i, _i64, _vp, _vp, _vp],
    "brawl_cuda_metropolis_replay": [_vp, _i, _d, _i64, _i, _vp, _vp],
    "brawl_cuda_metropolis_replay_sampled": [_vp, _i, _d, _i64, _i64, _i, _vp, _vp, _vp],
    "brawl_cuda_metropolis_run": [_vp, _vp, _i64, _i, _u64, _u64, _vp, _vp, _vp, _vp],
    "brawl_cuda_metropolis_enqueue": [_vp, _vp, _i64, _i, _u64, _u64, _vp, _vp, _vp],
    "brawl_cuda_metropolis_counters": [_vp, _i, _vp, _vp, _vp],
    "brawl_cuda_metropolis_last_launches": [_vp, _vp],
    "brawl_cuda_metropolis_tune": [_vp, _i, _i, _i, _i],
    "brawl_cuda_metropolis_set_mode": [_vp, _i],
    "brawl_cuda_metropolis_set_layout": [_vp, _i],
    "brawl_cuda_metropolis_plan": [_vp, _i, _vp],
    "brawl_cuda_radial_counts": [_vp, _i, _i, _vp, _vp],
    "brawl_cuda_radial_counts_batch": [_vp, _i, _i, _i, _vp, _vp],
    "brawl_cuda_wl_sweeps_replay": [_vp, _i, _vp, _vp, _vp, _i, _i, _i, _d, _i64, _i, _vp, _vp, _vp],
    "brawl_cuda_wl_sweeps": [_vp, _i, _vp, _vp, _i, _vp, _i, _vp, _vp, _i, _d, _i64, _i, _u64, _u64, _vp, _vp],
    "brawl_cuda_wl_enter_window": [_vp, _i, _vp, _vp, _vp, _d, _i64, _u64, _u64, _vp, _vp],
    "brawl_cuda_wl_enter_window_replay": [_vp, _i, _d, _d, _d, _d, _d, _i64, _vp, _i64, _i, _vp, _vp, _vp, _vp],
    "brawl_cuda_lattice_ptr": [_vp, _vp, _vp],
    "brawl_cuda_swap_replicas": [_vp, _i, _i],
    "brawl_cuda_wl_window_average": [_vp, _vp, _i, _i, _i, _d],
    "brawl_cuda_wl_init": [_vp, _i, _vp, _i],
    "brawl_cuda_wl_set_windows": [_vp, _vp, _vp, _i],
    "brawl_cuda_wl_zero_hist": [_vp],
    "brawl_cuda_wl_set_lng": [_vp, _vp],
    "brawl_cuda_wl_set_span": [_vp, _i],
    "brawl_cuda_wl_get": [_vp, _i, _vp],
    "brawl_cuda_wl_iterate": [_vp, _d, _i64, _i, _u64, _u64, _vp, _vp, _vp, _vp],
    "brawl_cuda_comm_unique_id": [_vp],
    "brawl_cuda_comm_create": [_vp, _i, _i, _vp],
    "brawl_cuda_comm_destroy": [_vp],
    "brawl_cuda_comm_allgather": [_vp, _vp, _i, _vp],
    "brawl_cuda_wl_allreduce": [_vp, _vp, _i],
    "brawl_cuda_wl_allgather_lng": [_vp, _vp],
    "brawl_cuda_exchange_replica": [_vp, _i, _i],
    "brawl_cuda_exchange_replicas": [_vp, _i, _vp, _vp],
    "brawl_cuda_swap_replicas_batch": [_vp, _i, _vp, _vp],
    "brawl_cuda_copy_replicas_batch": [_vp, _i, _vp, _vp],
    "brawl_cuda_ns_walk_replay": [_vp, _i, _vp, _d, _i64, _vp, _vp],
    "brawl_cuda_ns_walk": [_vp, _i, _vp, _vp, _vp, _i64, _u64, _u64, _vp],
}
EXPORTS = sorted(list(_SIGNATURES) + ["brawl_cuda_last_error"])


def load():
    """dlopen the library and attach argtypes.  Does not touch the GPU."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise BrawlCudaError(
                "%s not found: build it with `python -m brawl_b200.build` (nvcc, sm_100a). "
                "brawl_b200 has no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.brawl_cuda_last_error.restype = C.c_char_p
        for name, args in _SIGNATURES.items():
            f = getattr(L, name)
            f.argtypes = args
            f.restype = C.c_int
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise BrawlCudaError(load().brawl_cuda_last_error().decode())
