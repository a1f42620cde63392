"""The reference's workload loop (generate_data.py:19-261) as batches on the GPU.

``generate(file_path, params, nbr_datapoints, fixed_errors)`` keeps the reference's signature, parameter dictionary
and on-disk format: a pandas DataFrame indexed by (data_nr, type) with one 'data' column -- row (-1, 0) holds
``params``, row (i, 0) the uint8 error lattice of data point i *before* its class was hidden, row (i, 1) the decoder's
class distribution (generate_data.py:232-255) -- so ``MCMCDataReader`` (src/mcmc.py) and downstream consumers read
GPU-generated files unchanged.

Per batch the device draws the errors (generate_random_error), labels them (define_equivalence_class -> eq_true),
hides the class (apply_random_logical), decodes, and counts ``argmax != eq_true`` (argmin for "ST"); lattices stay in
HBM between those steps for the decoders that have a device-pointer entry (STDC, PTEQ family), the others take the
hidden lattices through their host-buffer ``*_batch`` entry.  ``params['mwpm_init']`` (generate_data.py:126-128, planar
only) starts every class's chains from its class-constrained minimum-weight matching (src/mwpm.py, matched on the host
by the native library); ``mwpm_batch`` scores the matching decoders themselves ("MWPM" / "eMWPM",
generate_data_noise_models.py:123-146).
"""
import numpy as np

from . import _lib
from . import decoders as _dec
from . import decoders_biasednoise as _decb
from .src import mcmc as _mcmc
from .src import mwpm as _mwpm
from .src.toric_model import Toric_code
from .src.planar_model import Planar_code
from .src.rotated_surface_model import RotSurCode
from .src.xzzx_model import xzzx_code

# cap on Ladder.step calls of the PTEQ family when params has no 'pt_steps': the reference's own default
# (decoders.py:25, decoders_biasednoise.py:28,93,175: steps=50000000).  The history the convergence criterion needs is
# 2 bytes per step per ladder; the driver splits a batch into waves that fit the device.
PT_STEPS = 50000000

_CODES = {'toric': Toric_code, 'planar': Planar_code, 'rotated': RotSurCode, 'xzzx': xzzx_code}


def noise_probabilities(params):
    """(p_error, None) for the toric form or (None, (p_x, p_y, p_z)) -- generate_data.py:57-118."""
    code, noise = params['code'], params.get('noise', 'depolarizing')
    if code == 'toric':
        assert noise == 'depolarizing'
        return params['p_error'], None
    if code == 'planar':
        assert noise in ['depolarizing', 'alpha']
    if noise == 'depolarizing':
        p = params['p_error'] / 3
        return None, (p, p, p)
    if noise == 'biased':
        eta, p = params['eta'], params['p_error']
        p_x = p / (2 * (eta + 1))
        return None, (p_x, p_x, p * eta / (eta + 1))
    if noise == 'alpha':
        pz_tilde, alpha = params['p_error'], params['alpha']
        p_tilde = pz_tilde + 2 * pz_tilde**alpha
        p = p_tilde / (1 + p_tilde)
        p_x = pz_tilde**alpha * (1 - p)
        return None, (p_x, p_x, pz_tilde * (1 - p))
    raise ValueError(f"unknown noise model {noise!r}")


def _wrap(code_cls, size, qm):
    out = []
    for q in qm:
        c = code_cls(size)
        c.qubit_matrix = q.reshape(c.qubit_matrix.shape).copy()
        out.append(c)
    return out


def mwpm_start_states(params, qubit):
    """generate_data.py:126-128 (`mwpm_init`): the per-class start chains class_sorted_mwpm gives for every error chain
    of the batch (numpy [S, n_sites]) -> [S, 4, n_sites], matched on the host (csrc/qecmc_mwpm.cu)."""
    assert params['code'] == 'planar'
    L = params['size']
    return _mwpm.class_sorted_mwpm_batch(qubit.reshape(-1, 2, L, L), L).reshape(qubit.shape[0], 4, -1)


def mwpm_batch(params, qubit, eq_true=None):
    """The matching decoders as a workload (generate_data_noise_models.py:123-146), host only: "MWPM" = class of the
    unconstrained matching (regular_mwpm), "eMWPM" = the class whose constrained matching is shortest (enhanced_mwpm,
    depolarizing rule; ties go to the lowest class here, the reference draws among them).
    -> dict(choice=[S], failures=int or None)"""
    assert params['code'] == 'planar'
    L = params['size']
    qm = np.ascontiguousarray(qubit, dtype=np.uint8).reshape(-1, 2, L, L)
    if params['method'] == 'MWPM':
        sol = _lib.mwpm_planar(L, qm=qm, class_sorted=False)[0]
        choice = _lib.host_planar_class(sol)
    elif params['method'] == 'eMWPM':
        chains = _lib.mwpm_planar(L, qm=qm, class_sorted=True)[0]
        choice = np.count_nonzero(chains.reshape(qm.shape[0], 4, -1), axis=2).argmin(axis=1)
    else:
        raise ValueError("mwpm_batch scores 'MWPM' or 'eMWPM'")
    choice = choice.astype(np.int32)
    return dict(choice=choice, failures=None if eq_true is None else int((choice != np.asarray(eq_true)).sum()))


def decode_batch(params, hidden, seed=None, device=0):
    """Decoder dispatch of generate_data.py:137-201 on a batch of hidden lattices (numpy [S, n_sites]), or of per-class
    start chains (numpy [S, n_eq, n_sites], the `mwpm_init` route).  -> (distributions [S, n_eq], use_argmin)"""
    method, noise = params['method'], params.get('noise', 'depolarizing')
    code_cls, size = _CODES[params['code']], params['size']
    codes = [_wrap(code_cls, size, h) for h in hidden] if hidden.ndim == 3 else _wrap(code_cls, size, hidden)
    kw = dict(seed=seed, device=device)
    if method == 'PTEQ':
        if noise == 'depolarizing':
            return _dec.PTEQ_batch(codes, params['p_error'], steps=params.get('pt_steps', PT_STEPS), **kw), False
        if noise == 'biased':
            p, eta = params['p_error'], params['eta']
            pz_tilde = (p / (1 + 1 / eta)) / (1 - p)
            alpha = np.log(pz_tilde / (2 * eta)) / np.log(pz_tilde)
            return _decb.PTEQ_alpha_batch(codes, pz_tilde, alpha=alpha, steps=params.get('pt_steps', PT_STEPS), **kw), False
        if noise == 'alpha':
            return _decb.PTEQ_alpha_batch(codes, params['p_error'], alpha=params['alpha'], Nc=params.get('Nc'),
                                          SEQ=params.get('SEQ', 2), TOPS=params.get('TOPS', 10), eps=params.get('eps', 0.1),
                                          iters=params.get('iters', 10), conv_criteria=params.get('conv_criteria', 'error_based'),
                                          steps=params.get('pt_steps', PT_STEPS), **kw), False
    if method == 'PTEQ_with_shortest':
        assert noise == 'alpha'
        out = _decb.PTEQ_alpha_with_shortest_batch(codes, params['p_error'], alpha=params['alpha'],
                                                   steps=params.get('pt_steps', PT_STEPS), **kw)
        return np.concatenate([np.asarray(o, dtype=np.float64) for o in out], axis=1), False
    if method == 'PTDC':
        return _dec.PTDC_batch(codes, params['p_error'], params['p_sampling'], **kw), False
    if method == 'PTRC':
        return _dec.PTRC_batch(codes, params['p_error'], params['p_sampling'], **kw), False
    if method == 'STDC':
        return _dec.STDC_batch(codes, params['p_error'], params['p_sampling'], steps=params['steps'],
                               droplets=params['droplets'], **kw), False
    if method == 'STDC_N_n':
        assert noise == 'alpha'
        return _dec.STDC_Nall_n_alpha_batch(codes, params['p_sampling'], params['alpha'], params['p_error'],
                                            steps=params['steps'], **kw), False
    if method == 'ST':
        return _dec.single_temp_batch(codes, params['p_error'], params['steps'], **kw), True
    if method == 'STRC':
        return _dec.STRC_batch(codes, params['p_error'], p_sampling=params['p_sampling'], steps=params['steps'],
                               droplets=params['droplets'], **kw), False
    if method in ('MWPM', 'eMWPM'):
        raise ValueError("MWPM / eMWPM score the matching itself (generate_data_noise_models.py:123-146): use mwpm_batch")
    raise ValueError(f"unknown method {method!r}")


def generate_batch(params, S, seed=0, device=0):
    """One batch of the workload loop -> dict(qubit=[S, n_sites] uint8, eq_true=[S], distr=[S, n_eq], choice=[S],
    failures=int).  The STDC and depolarizing-PTEQ paths keep the lattices on the device from error generation to
    failure counting; the others stage the hidden lattices through the host-buffer decoder entry."""
    import torch
    mwpm_init = bool(params.get('mwpm_init'))
    ctx = _lib.default_context(device)
    geom = _lib.GEOM_NAMES[params['code']]
    L = params['size']
    n, n_eq = _lib.nsites(geom, L), _lib.neq(geom)
    p_error, p_xyz = noise_probabilities(params)
    dev = torch.device('cuda', device)
    with torch.cuda.device(dev):
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        d_q = torch.empty((S, n), dtype=torch.uint8, device=dev)
        d_true = torch.empty(S, dtype=torch.int32, device=dev)
        ctx.generate_errors_dev(geom, L, S, d_q.data_ptr(), d_true.data_ptr(), p_error=p_error, p_xyz=p_xyz, seed=seed)
        d_hidden = d_q.clone()                                                    # df_qubit keeps the unhidden error
        ctx.apply_random_logical_dev(geom, L, d_hidden.data_ptr(), S, seed=seed ^ 0x5A5A5A5A)
        method, noise = params['method'], params.get('noise', 'depolarizing')
        use_argmin = False
        d_choice = torch.empty(S, dtype=torch.int32, device=dev)
        if mwpm_init:
            # generate_data.py:126-128: start every class's chains from its minimum-weight matching instead of the
            # hidden error (no apply_random_logical: the matchings carry no trace of the error's class)
            starts = mwpm_start_states(params, d_q.cpu().numpy())
            distr, use_argmin = decode_batch(params, starts, seed=seed + 1, device=device)
            ctx.set_stream(torch.cuda.current_stream().cuda_stream)
            scored = np.ascontiguousarray(distr[:, :n_eq])
            d_scored = torch.from_numpy(scored).to(dev)
            failures = ctx.count_failures_dev(d_scored.data_ptr(), _lib.DISTR_U8 if scored.dtype == np.uint8 else _lib.DISTR_F64,
                                              n_eq, S, d_true.data_ptr(), d_choice.data_ptr(), use_argmin=use_argmin)
        elif method == 'STDC' and geom in (_lib.TORIC, _lib.PLANAR):
            d_out = torch.empty((S, n_eq), dtype=torch.float64, device=dev)
            code0 = _CODES[params['code']](L)
            ctx.stdc_dev(geom, _mcmc.fast_path_geometry(code0), L, d_hidden.data_ptr(), S, d_out.data_ptr(), params['p_error'],
                         params['p_sampling'] or params['p_error'], int(params['droplets']), int(params['steps']), iters=5,
                         per_class=False, randomize=True, seed=seed + 1, want_stats=False)
            failures = ctx.count_failures_dev(d_out.data_ptr(), _lib.DISTR_F64, n_eq, S, d_true.data_ptr(), d_choice.data_ptr())
            distr = d_out.cpu().numpy()
        elif method == 'PTEQ' and noise == 'depolarizing':
            d_out = torch.empty((S, n_eq), dtype=torch.uint8, device=dev)
            ctx.pteq_dev(geom, L, _lib.LADDER_DEPOLARIZING, d_hidden.data_ptr(), S, d_out.data_ptr(), params['p_error'],
                         Nc=params.get('Nc'), SEQ=params.get('SEQ', 2), TOPS=params.get('TOPS', 10), eps=params.get('eps', 0.1),
                         steps=params.get('pt_steps', PT_STEPS), iters=params.get('iters', 10), seed=seed + 1)
            failures = ctx.count_failures_dev(d_out.data_ptr(), _lib.DISTR_U8, n_eq, S, d_true.data_ptr(), d_choice.data_ptr())
            distr = d_out.cpu().numpy()
        else:
            distr, use_argmin = decode_batch(params, d_hidden.cpu().numpy(), seed=seed + 1, device=device)
            ctx.set_stream(torch.cuda.current_stream().cuda_stream)
            # PTEQ_with_shortest scores on the first four entries only (generate_data.py:170)
            scored = np.ascontiguousarray(distr[:, :n_eq])
            d_scored = torch.from_numpy(scored).to(dev)
            failures = ctx.count_failures_dev(d_scored.data_ptr(), _lib.DISTR_U8 if scored.dtype == np.uint8 else _lib.DISTR_F64,
                                              n_eq, S, d_true.data_ptr(), d_choice.data_ptr(), use_argmin=use_argmin)
        out = dict(qubit=d_q.cpu().numpy(), eq_true=d_true.cpu().numpy(), distr=distr, choice=d_choice.cpu().numpy(),
                   failures=int(failures))
        ctx.set_stream(None)
    return out


def auto_batch(params, device=0):
    """Syndromes per batch that fill the GPU for this workload: for the STDC family (STDC, STRC, ST) a whole number of rounds
    of the chain kernel's CTAs within one wave of the table budget, read from the library's plan of the most recent call
    (qecmc_last_plan); for the tempering decoders one resident ladder per slot of the tempering kernel's grid."""
    ctx = _lib.default_context(device)
    sms = ctx.device_info()["sm_count"]
    if params['method'] in ('STDC', 'STRC', 'ST'):
        wave_cap, round_chains = ctx.last_plan()
        chains = _lib.neq(_lib.GEOM_NAMES[params['code']]) * (1 if params['method'] == 'ST' else int(params['droplets']))
        per_round = max(1, round_chains // chains)
        if wave_cap <= 0 or round_chains <= 0:
            return 256
        return int(max(64, min(wave_cap // per_round * per_round if wave_cap >= per_round else wave_cap, 8 * per_round)))
    return 32 * sms


def generate(file_path, params, nbr_datapoints=10**6, fixed_errors=None, batch=None, seed=None, device=0, verbose=True):
    """generate_data.py:19-261.  Extra keyword arguments (batch, seed, device) have no reference counterpart: the
    reference decodes one syndrome at a time, unseeded.  batch = None sizes the batches itself: a first batch of 256, then
    whatever fills the GPU for this workload (auto_batch) -- a planar d = 7 round holds 49 728 syndromes, and batches of 256
    would leave the device idle."""
    auto = batch is None
    if auto:
        batch = 256
    import pandas as pd
    names = ['data_nr', 'type']
    frames = [pd.DataFrame([[params]], index=pd.MultiIndex.from_product([[-1], np.arange(1)], names=names), columns=['data'])]
    if verbose:
        print('\nDataFrame with opened at: ' + str(file_path))
    if fixed_errors is not None:
        nbr_datapoints = 10000000
    failed_syndroms = 0
    seed = _dec._next_seed(seed)
    shape = (2, params['size'], params['size']) if params['code'] in ('toric', 'planar') else (params['size'], params['size'])
    i = 0
    while i < nbr_datapoints:
        S = min(batch, nbr_datapoints - i)
        res = generate_batch(params, S, seed=seed + 7919 * i, device=device)
        rows, idx = [], []
        for s in range(S):
            rows.append([res['qubit'][s].reshape(shape).astype(np.uint8)])
            idx.append((i + s, 0))
            rows.append([np.array(res['distr'][s])])
            idx.append((i + s, 1))
            if params['method'] != 'STDC_N_n':        # generate_data.py:188-195 never scores STDC_N_n
                failed_syndroms += int(res['choice'][s] != res['eq_true'][s])
            if fixed_errors is not None and failed_syndroms == fixed_errors:
                S = s + 1
                break
        if auto and i == 0:
            batch = auto_batch(params, device)
            if fixed_errors is not None:        # a batch is decoded whole: do not overshoot the failure target by much
                seen = max(failed_syndroms, 1) / float(S)
                batch = int(max(256, min(batch, 1.5 * fixed_errors / seen)))
        rows, idx = rows[:2 * S], idx[:2 * S]
        frames.append(pd.DataFrame(rows, index=pd.MultiIndex.from_tuples(idx, names=names), columns=['data']))
        i += S
        df = pd.concat(frames)
        df.to_pickle(file_path)                       # intermediate save point (writing over), generate_data.py:244-248
        if verbose:
            print('Failed so far:', failed_syndroms, 'of', i)
        if fixed_errors is not None and failed_syndroms == fixed_errors:
            if verbose:
                print('Desired amount of failes syndroms achieved, breaking loop.')
            break
    df = pd.concat(frames)
    df.to_pickle(file_path)
    if verbose:
        print('\nCompleted')
    return failed_syndroms, i
