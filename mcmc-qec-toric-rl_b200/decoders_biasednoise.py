"""Biased / alpha-noise parallel-tempering decoders with the reference's signatures (decoders_biasednoise.py)."""
from . import _lib
from .decoders import PTEQ_batch


def PTEQ_biased_batch(init_codes, p, eta=0.5, Nc=None, SEQ=2, TOPS=10, tops_burn=2, eps=0.1, steps=50000000, iters=10,
                      conv_criteria='error_based', seed=None, device=0, return_info=False):
    """PTEQ_biased (decoders_biasednoise.py:28-75) over a batch -> uint8 [S, nbr_eq_classes]."""
    return PTEQ_batch(init_codes, p, Nc, SEQ, TOPS, tops_burn, eps, steps, iters, conv_criteria, seed, device, return_info,
                      _kind=_lib.LADDER_BIASED, _param_b=float(eta))


def PTEQ_biased(init_code, p, eta=0.5, Nc=None, SEQ=2, TOPS=10, tops_burn=2, eps=0.1, steps=50000000, iters=10,
                conv_criteria='error_based'):
    return PTEQ_biased_batch([init_code], p, eta, Nc, SEQ, TOPS, tops_burn, eps, steps, iters, conv_criteria)[0]


def PTEQ_alpha_batch(init_codes, pz_tilde, alpha=1, Nc=None, SEQ=2, TOPS=10, tops_burn=2, eps=0.1, steps=50000000, iters=10,
                     conv_criteria='error_based', seed=None, device=0, return_info=False):
    """PTEQ_alpha (decoders_biasednoise.py:175-222) over a batch -> uint8 [S, nbr_eq_classes]."""
    return PTEQ_batch(init_codes, pz_tilde, Nc, SEQ, TOPS, tops_burn, eps, steps, iters, conv_criteria, seed, device,
                      return_info, _kind=_lib.LADDER_ALPHA, _param_b=float(alpha))


def PTEQ_alpha(init_code, pz_tilde, alpha=1, Nc=None, SEQ=2, TOPS=10, tops_burn=2, eps=0.1, steps=50000000, iters=10,
               conv_criteria='error_based'):
    return PTEQ_alpha_batch([init_code], pz_tilde, alpha, Nc, SEQ, TOPS, tops_burn, eps, steps, iters, conv_criteria)[0]


def PTEQ_alpha_with_shortest(*a, **k):
    raise NotImplementedError("PTEQ_alpha_with_shortest (decoders_biasednoise.py:93-172) is not implemented on the device "
                              "path yet (DESIGN.md section 8); there is no CPU fallback")
