"""ctypes binding of include/lattigpu.h (the C-ABI drop-in boundary).

The shared library is built in-tree by lattigo-fhe-by-go_b200/build.py.  There is
no CPU fallback: if the library is missing or no CUDA device is present every
entry point raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LATTIGPU_LIB") or os.path.join(os.path.dirname(_HERE), "lib", "liblattigpu.so")
HEADER_PATH = os.path.join(os.path.dirname(os.path.dirname(_HERE)), "include", "lattigpu.h")

u64 = C.c_uint64
p64 = C.POINTER(C.c_uint64)
vp = C.c_void_p
ci = C.c_int


class LattigpuError(RuntimeError):
    pass


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LattigpuError(
                "liblattigpu.so not built (run `python lattigo-fhe-by-go_b200/build.py`); there is no CPU fallback")
        _lib = C.CDLL(LIB_PATH)
        _declare(_lib)
    return _lib


def check(rc):
    if rc != 0:
        raise LattigpuError("lattigpu error %d: %s" % (rc, lib().lg_last_error().decode()))


# (name, restype, argtypes); every int-returning entry is status-checked by the wrappers
_R, _P = vp, vp  # ring / poly handles are opaque pointers
_OP3 = [_R, ci, _P, _P, _P, vp]
_OP2 = [_R, ci, _P, _P, vp]
SIGNATURES = {
    "lg_last_error": (C.c_char_p, []),
    "lg_version": (C.c_char_p, []),
    "lg_device_count": (ci, [C.POINTER(ci)]),
    "lg_set_device": (ci, [ci]),
    "lg_stream_create": (ci, [C.POINTER(vp)]),
    "lg_stream_destroy": (ci, [vp]),
    "lg_stream_sync": (ci, [vp]),
    "lg_launch_count": (u64, []),
    "lg_debug_set_switch": (ci, [C.c_char_p, u64]),
    "lg_ring_create": (ci, [u64, ci, p64, C.POINTER(vp)]),
    "lg_ring_create_from_tables": (ci, [u64, ci, p64, p64, p64, p64, p64, p64, p64, C.POINTER(vp)]),
    "lg_ring_destroy": (ci, [_R]),
    "lg_ring_n": (u64, [_R]),
    "lg_ring_nlimbs": (ci, [_R]),
    "lg_ring_get_tables": (ci, [_R, p64, p64, p64, p64, p64, p64, p64]),
    "lg_is_prime": (ci, [u64]),
    "lg_generate_ntt_primes": (ci, [u64, u64, u64, p64]),
    "lg_primitive_root": (u64, [u64]),
    "lg_poly_create": (ci, [u64, ci, ci, C.POINTER(vp)]),
    "lg_poly_wrap": (ci, [vp, u64, ci, ci, C.POINTER(vp)]),
    "lg_poly_view": (ci, [_P, ci, ci, C.POINTER(vp)]),
    "lg_poly_destroy": (ci, [_P]),
    "lg_poly_get_data_len": (u64, [_P, ci, ci]),
    "lg_poly_write_to": (ci, [_P, ci, ci, C.c_char_p, u64, ci, vp]),
    "lg_poly_decode": (ci, [_P, ci, C.c_char_p, u64, ci, ci, vp]),
    "lg_poly_n": (u64, [_P]),
    "lg_poly_nlimbs": (ci, [_P]),
    "lg_poly_batch": (ci, [_P]),
    "lg_poly_device_ptr": (vp, [_P]),
    "lg_poly_batch_stride": (C.c_size_t, [_P]),
    "lg_poly_upload": (ci, [_P, ci, ci, ci, ci, p64, vp]),
    "lg_poly_download": (ci, [_P, ci, ci, ci, ci, p64, vp]),
    "lg_poly_upload_async": (ci, [_P, ci, ci, ci, ci, p64, vp]),
    "lg_poly_download_async": (ci, [_P, ci, ci, ci, ci, p64, vp]),
    "lg_poly_zero": (ci, [_P, vp]),
    "lg_poly_copy": (ci, [_P, ci, _P, vp]),
    "lg_ring_ntt": (ci, _OP2),
    "lg_ring_invntt": (ci, _OP2),
    "lg_ring_ntt_limb": (ci, [_R, ci, _P, ci, _P, ci, vp]),
    "lg_ring_invntt_limb": (ci, [_R, ci, _P, ci, _P, ci, vp]),
    "lg_ring_add": (ci, _OP3),
    "lg_ring_add_nomod": (ci, _OP3),
    "lg_ring_sub": (ci, _OP3),
    "lg_ring_sub_nomod": (ci, _OP3),
    "lg_ring_neg": (ci, _OP2),
    "lg_ring_reduce": (ci, _OP2),
    "lg_ring_mod": (ci, [_R, ci, _P, u64, _P, vp]),
    "lg_ring_and": (ci, [_R, ci, _P, u64, _P, vp]),
    "lg_ring_or": (ci, [_R, ci, _P, u64, _P, vp]),
    "lg_ring_xor": (ci, [_R, ci, _P, u64, _P, vp]),
    "lg_ring_mul_coeffs": (ci, _OP3),
    "lg_ring_mul_coeffs_and_add": (ci, _OP3),
    "lg_ring_mul_coeffs_and_add_nomod": (ci, _OP3),
    "lg_ring_mul_coeffs_constant": (ci, _OP3),
    "lg_ring_mul_coeffs_montgomery": (ci, _OP3),
    "lg_ring_mul_coeffs_montgomery_and_add": (ci, _OP3),
    "lg_ring_mul_coeffs_montgomery_and_add_nomod": (ci, _OP3),
    "lg_ring_mul_coeffs_montgomery_constant_and_add_nomod": (ci, _OP3),
    "lg_ring_mul_coeffs_montgomery_and_sub": (ci, _OP3),
    "lg_ring_mul_coeffs_montgomery_and_sub_nomod": (ci, _OP3),
    "lg_ring_mul_coeffs_montgomery_constant": (ci, _OP3),
    "lg_ring_mform": (ci, _OP2),
    "lg_ring_invmform": (ci, _OP2),
    "lg_ring_add_scalar": (ci, [_R, ci, _P, p64, vp]),
    "lg_ring_sub_scalar": (ci, [_R, ci, _P, p64, vp]),
    "lg_ring_mul_scalar": (ci, [_R, ci, _P, p64, _P, vp]),
    "lg_ring_add_scalar_halves": (ci, [_R, ci, _P, p64, p64, _P, vp]),
    "lg_ring_mul_scalar_montgomery_halves": (ci, [_R, ci, _P, p64, p64, _P, vp]),
    "lg_ring_mul_scalar_montgomery_halves_and_add": (ci, [_R, ci, _P, p64, p64, _P, vp]),
    "lg_ring_mul_by_pow2": (ci, [_R, ci, _P, u64, _P, vp]),
    "lg_ring_mult_by_monomial": (ci, [_R, ci, _P, u64, _P, vp]),
    "lg_ring_mul_by_vector_montgomery": (ci, [_R, ci, _P, _P, _P, vp]),
    "lg_ring_mul_by_vector_montgomery_and_add_nomod": (ci, [_R, ci, _P, _P, _P, vp]),
    "lg_ring_bitreverse": (ci, _OP2),
    "lg_ring_mul_poly": (ci, [_R, _P, _P, _P, ci, vp]),
    "lg_ring_mul_poly_naive": (ci, [_R, _P, _P, _P, ci, vp]),
    "lg_ring_exp": (ci, [_R, _P, u64, _P, vp]),
    "lg_ring_shift": (ci, [_R, _P, u64, _P, vp]),
    "lg_ring_rotate": (ci, [_R, _P, u64, vp]),
    "lg_ring_equal": (ci, [_R, ci, _P, _P, C.POINTER(ci), vp]),
    "lg_galois_create": (ci, [u64, u64, u64, C.POINTER(vp)]),
    "lg_galois_create_from_index": (ci, [p64, u64, C.POINTER(vp)]),
    "lg_galois_get_index": (ci, [vp, p64]),
    "lg_galois_destroy": (ci, [vp]),
    "lg_ring_permute_ntt_with_index": (ci, [ci, _P, vp, _P, vp]),
    "lg_ring_permute_ntt": (ci, [ci, _P, u64, _P, vp]),
    "lg_ring_permute": (ci, [_R, ci, _P, u64, _P, vp]),
    "lg_ring_div_floor_by_last_modulus_ntt": (ci, [_R, ci, _P, vp]),
    "lg_ring_div_floor_by_last_modulus": (ci, [_R, ci, _P, vp]),
    "lg_ring_div_floor_by_last_modulus_many_ntt": (ci, [_R, ci, _P, ci, vp]),
    "lg_ring_div_floor_by_last_modulus_many": (ci, [_R, ci, _P, ci, vp]),
    "lg_ring_div_round_by_last_modulus_ntt": (ci, [_R, ci, _P, vp]),
    "lg_ring_div_round_by_last_modulus": (ci, [_R, ci, _P, vp]),
    "lg_ring_div_round_by_last_modulus_many_ntt": (ci, [_R, ci, _P, ci, vp]),
    "lg_ring_div_round_by_last_modulus_many": (ci, [_R, ci, _P, ci, vp]),
    "lg_extender_create": (ci, [_R, _R, C.POINTER(vp)]),
    "lg_extender_destroy": (ci, [vp]),
    "lg_extender_modup_split_qp": (ci, [vp, ci, _P, _P, vp]),
    "lg_extender_modup_split_pq": (ci, [vp, ci, _P, _P, vp]),
    "lg_extender_moddown_ntt_pq": (ci, [vp, ci, _P, _P, vp]),
    "lg_extender_moddown_splited_ntt_pq": (ci, [vp, ci, _P, _P, _P, vp]),
    "lg_extender_moddown_pq": (ci, [vp, ci, _P, _P, vp]),
    "lg_extender_moddown_splited_pq": (ci, [vp, ci, _P, _P, _P, vp]),
    "lg_extender_moddown_splited_qp": (ci, [vp, ci, ci, _P, _P, _P, vp]),
    "lg_decomposer_create": (ci, [u64, p64, ci, p64, ci, C.POINTER(vp)]),
    "lg_decomposer_destroy": (ci, [vp]),
    "lg_decomposer_beta": (ci, [vp]),
    "lg_decomposer_xalpha": (ci, [vp, ci]),
    "lg_decomposer_decompose": (ci, [vp, ci, ci, _P, _P, vp]),
    "lg_decomposer_decompose_and_split": (ci, [vp, ci, ci, _P, _P, _P, vp]),
    "lg_ckks_eval_create": (ci, [_R, _R, C.POINTER(vp)]),
    "lg_ckks_eval_destroy": (ci, [vp]),
    "lg_swk_create": (ci, [u64, ci, ci, p64, C.POINTER(vp)]),
    "lg_swk_alloc": (ci, [u64, ci, ci, C.POINTER(vp)]),
    "lg_swk_poly": (ci, [vp, ci, ci, C.POINTER(vp)]),
    "lg_swk_invalidate": (ci, [vp]),
    "lg_swk_beta": (ci, [vp]),
    "lg_swk_nlimbs": (ci, [vp]),
    "lg_swk_n": (u64, [vp]),
    "lg_swk_wrap": (ci, [vp, u64, ci, ci, C.POINTER(vp)]),
    "lg_swk_destroy": (ci, [vp]),
    "lg_ckks_switch_keys_in_place": (ci, [vp, ci, _P, vp, _P, _P, vp]),
    "lg_ckks_mul_relin": (ci, [vp, ci, _P, _P, _P, _P, vp, _P, _P, vp]),
    "lg_ckks_relinearize": (ci, [vp, ci, _P, _P, _P, vp, _P, _P, vp]),
    "lg_ckks_rescale": (ci, [vp, ci, _P, _P, ci, vp]),
    "lg_ckks_switch_keys": (ci, [vp, ci, _P, _P, vp, _P, _P, vp]),
    "lg_ckks_permute_ntt": (ci, [vp, ci, _P, _P, vp, vp, _P, _P, vp]),
    "lg_ckks_hoist": (ci, [vp, ci, _P, C.POINTER(vp), vp]),
    "lg_ckks_switch_key_hoisted": (ci, [vp, vp, _P, vp, vp, _P, _P, vp]),
    "lg_hoisted_destroy": (ci, [vp]),
    "lg_bfv_eval_create": (ci, [_R, _R, _R, u64, C.POINTER(vp)]),
    "lg_bfv_eval_destroy": (ci, [vp]),
    "lg_bfv_mul": (ci, [vp, _P, _P, _P, _P, _P, _P, _P, vp]),
    "lg_bfv_switch_keys_core": (ci, [vp, _P, vp, _P, _P, vp]),
    "lg_bfv_relinearize": (ci, [vp, _P, _P, _P, vp, _P, _P, vp]),
    "lg_bfv_switch_keys": (ci, [vp, _P, _P, vp, _P, _P, vp]),
    "lg_bfv_permute": (ci, [vp, _P, _P, u64, vp, _P, _P, vp]),
    "lg_comm_get_unique_id": (ci, [C.POINTER(C.c_uint8)]),
    "lg_comm_create": (ci, [ci, ci, C.POINTER(C.c_uint8), C.POINTER(vp)]),
    "lg_comm_destroy": (ci, [vp]),
    "lg_comm_world": (ci, [vp]),
    "lg_comm_rank": (ci, [vp]),
    "lg_comm_aggregate_shares": (ci, [vp, _R, ci, _P, vp]),
    "lg_comm_limb_owner": (ci, [ci, ci]),
    "lg_comm_xbuf_words_needed": (C.c_size_t, [u64, ci, ci, ci]),
    "lg_comm_xbuf_alloc": (ci, [vp, C.c_size_t, C.POINTER(C.c_uint8)]),
    "lg_comm_xbuf_open": (ci, [vp, ci, C.POINTER(C.c_uint8)]),
    "lg_comm_xbuf_attach": (ci, [vp, ci, vp]),
    "lg_comm_xbuf_words": (C.c_size_t, [vp]),
    "lg_comm_check": (ci, [vp, vp]),
    "lg_comm_gather_limbs": (ci, [vp, _R, ci, _P, vp]),
    "lg_ckks_switch_keys_in_place_resident": (ci, [vp, vp, ci, _P, vp, _P, _P, vp]),
    "lg_ckks_mul_relin_rescale_resident": (ci, [vp, vp, ci, _P, _P, _P, _P, vp, _P, _P, ci, vp]),
    "lg_ckks_rescale_resident": (ci, [vp, vp, ci, _P, _P, vp]),
    "lg_ckks_switch_keys_in_place_sharded": (ci, [vp, vp, ci, _P, vp, _P, _P, vp]),
    "lg_ckks_mul_relin_sharded": (ci, [vp, vp, ci, _P, _P, _P, _P, vp, _P, _P, vp]),
    "lg_ckks_rescale_sharded": (ci, [vp, vp, ci, _P, _P, vp]),
    "lg_scaler_create": (ci, [u64, _R, C.POINTER(vp)]),
    "lg_scaler_params_host": (ci, [u64, p64, ci, p64, C.POINTER(C.c_double), p64, p64]),
    "lg_scaler_destroy": (ci, [vp]),
    "lg_scaler_get_params": (ci, [vp, p64, C.POINTER(C.c_double)]),
    "lg_scaler_scale": (ci, [vp, _P, _P, vp]),
    "lg_bfv_lift_create": (ci, [_R, u64, C.POINTER(vp)]),
    "lg_bfv_lift_params_host": (ci, [p64, ci, u64, p64]),
    "lg_bfv_lift_destroy": (ci, [vp]),
    "lg_bfv_lift_get_params": (ci, [vp, p64]),
    "lg_bfv_lift_apply": (ci, [vp, _P, _P, vp]),
    "lg_prng_create": (ci, [C.c_char_p, C.c_size_t, C.POINTER(vp)]),
    "lg_prng_destroy": (ci, [vp]),
    "lg_prng_seed": (ci, [vp, C.c_char_p, C.c_size_t]),
    "lg_prng_get_clock": (u64, [vp]),
    "lg_prng_clock": (ci, [vp, C.c_char_p]),
    "lg_prng_set_clock": (ci, [vp, u64]),
    "lg_crp_create": (ci, [C.c_char_p, C.c_size_t, _R, C.POINTER(vp)]),
    "lg_crp_destroy": (ci, [vp]),
    "lg_crp_seed": (ci, [vp, C.c_char_p, C.c_size_t]),
    "lg_crp_get_clock": (u64, [vp]),
    "lg_crp_set_clock": (ci, [vp, u64]),
    "lg_crp_clock": (ci, [vp, _P, ci, vp]),
    "lg_crp_clock_host": (ci, [vp, p64]),
}


def _declare(L):
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args


def header_symbols():
    """names of every function declared in include/lattigpu.h"""
    import re

    with open(HEADER_PATH) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lg_[a-z0-9_]+)\s*\(", src)))
