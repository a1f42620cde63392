"""Biased / alpha-noise parallel-tempering decoders with the reference's signatures (decoders_biasednoise.py)."""
from . import _lib
from .decoders import PTEQ_batch, conv_crit_error_based_PT

# the biased and alpha ladders use the same test on their own history (decoders_biasednoise.py:79-90, 226-237)
conv_crit_error_based_PT_biased = conv_crit_error_based_PT
conv_crit_error_based_PT_alpha = conv_crit_error_based_PT


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


def PTEQ_alpha_with_shortest_batch(init_codes, pz_tilde, alpha=1, Nc=None, SEQ=2, TOPS=10, tops_burn=2, eps=0.1,
                                   steps=50000000, iters=10, conv_criteria='error_based', seed=None, device=0):
    """PTEQ_alpha_with_shortest (decoders_biasednoise.py:93-172) over a batch.  Returns the reference's three arrays,
    one row per syndrome: class percentages (uint8), the distribution from the distinct shortest chains
    (float64 percent) and the share of samples at the shortest effective length (float64 percent)."""
    import numpy as np
    from .decoders import _batch, _next_seed
    code, qm, per_class = _batch(init_codes)
    if per_class:
        raise TypeError("PTEQ_alpha_with_shortest takes a single code object per syndrome")
    Nc = Nc or code.system_size
    if tops_burn >= TOPS:
        print('tops_burn has to be smaller than TOPS')
    pct, slen, sn, su, info = _lib.default_context(device).pteq_shortest(
        code.geometry, code.system_size, _lib.LADDER_ALPHA, qm, pz_tilde, Nc=Nc, param_b=float(alpha), SEQ=SEQ, TOPS=TOPS,
        tops_burn=tops_burn, eps=eps, steps=int(steps), iters=int(iters), conv=(conv_criteria == 'error_based'), p_logical=0.5,
        seed=_next_seed(seed))
    beta = -np.log(pz_tilde)
    z = su * np.exp(-beta * slen)                    # every stored chain has the class's shortest effective length
    with np.errstate(invalid="ignore", divide="ignore"):
        return pct, z / z.sum(1, keepdims=True) * 100, sn / sn.sum(1, keepdims=True) * 100


def PTEQ_alpha_with_shortest(init_code, pz_tilde, alpha=1, Nc=None, SEQ=2, TOPS=10, tops_burn=2, eps=0.1, steps=50000000,
                             iters=10, conv_criteria='error_based'):
    a, b, c = PTEQ_alpha_with_shortest_batch([init_code], pz_tilde, alpha, Nc, SEQ, TOPS, tops_burn, eps, steps, iters,
                                             conv_criteria)
    return a[0], b[0], c[0]
