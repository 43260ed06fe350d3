"""CPU emulation of the tensor core's truncating fp32 accumulation inside the recurrent forward (H = 256 chain), to
evaluate accumulation schemes without GPU time.  z = corr (exact, added round-to-nearest) + main, main = chain(s) of
K=8 MMAs each doing acc = trunc_toward_zero_fp32(acc + exact_dot8).  Reports the parameter-gradient error (exact fp64
BPTT on the emulated tapes) and the persistence of the one-step error.
    python tools/sim_trunc.py [case] scheme...     schemes: exact chain36 split2 split4 rot rot_split2 biasout ..."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _util import cudalstm_weights_np, golden  # noqa: E402


def trunc32(v):
    t = v.astype(np.float32)
    over = np.abs(t.astype(np.float64)) > np.abs(v)
    t[over] = np.nextafter(t[over], np.float32(0))
    return t.astype(np.float64)


def rn32(v):
    return v.astype(np.float32).astype(np.float64)


def tf32_hi(v):     # round-to-nearest (ties away) to 10 mantissa bits, like split_tf32
    u = v.astype(np.float32).view(np.uint32)
    return ((u + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32).astype(np.float64)


def sigmoid(z):
    return 1.0 / (1.0 + np.exp(-z))


def run(case, scheme, comp):
    g = golden(case)
    B, T, I, H = [int(v) for v in g["meta"]][:4]
    w = cudalstm_weights_np(g)
    x = g["x_lstm"].astype(np.float64)
    W = np.concatenate([w["lstm.weight_ih_l0"], (w["lstm.bias_ih_l0"] + w["lstm.bias_hh_l0"])[:, None],
                        np.zeros((4 * H, (-(I + 1)) % 8)), w["lstm.weight_hh_l0"]], axis=1).astype(np.float64)   # [4H, K]
    KX = I + 1 + (-(I + 1)) % 8
    K = W.shape[1]
    bias_out = "biasout" in scheme
    bvec = W[:, I].copy()
    if bias_out:
        W[:, I] = 0.0
    Wc = rn32(W * (1.0 + comp))
    W_hi = tf32_hi(Wc)
    NJ = K // 8
    nsplit = 1
    for p_ in scheme:
        if p_.startswith("split"):
            nsplit = int(p_[5:])
    rot = "rot" in scheme
    h = np.zeros((B, H)); c = np.zeros((B, H))
    Hs, Cs, Gs = [], [], []
    for t in range(T):
        a = np.concatenate([x[:, t], np.ones((B, 1)), np.zeros((B, KX - I - 1)), h], axis=1)            # [B, K]
        a = rn32(a)
        if scheme[0] == "exact":
            z = rn32(a @ rn32(W).T) if not bias_out else rn32(rn32(a @ rn32(W).T) + bvec)
        else:
            a_hi = tf32_hi(a)
            exact = a @ Wc.T
            main_exact = a_hi @ W_hi.T
            corr = rn32(exact - main_exact)                      # correction terms: small accumulator, no loss
            order = list(range(NJ))
            if rot:
                r = (t * 7) % NJ
                order = order[r:] + order[:r]
                if (t // NJ) % 2:
                    order = order[::-1]
            if "xlast" in scheme:                   # input (+ bias) K steps after the recurrent ones
                order = order[KX // 8:] + order[:KX // 8]
            if "xsep" in scheme:                    # input (+ bias) K steps in their own accumulator
                parts = [np.array(order[:KX // 8])] + list(np.array_split(np.array(order[KX // 8:]), nsplit))
            else:
                parts = np.array_split(np.array(order), nsplit)
            z = corr.copy()
            off = None
            for p_ in scheme:
                if p_.startswith("offset"):       # offsetA_B: accumulators start at 1.5*A (i, g, o) / 1.5*B (f): fixed binade
                    oa, ob = [float(v) for v in p_[6:].split("_")]
                    off = np.full(4 * H, 1.5 * oa); off[H:2 * H] = 1.5 * ob
                    ulp = np.full(4 * H, oa * 2.0 ** -23); ulp[H:2 * H] = ob * 2.0 ** -23
            for part in parts:
                acc = np.zeros((B, 4 * H)) if off is None else np.tile(off, (B, 1))
                if off is not None:
                    z = z + len(part) * ulp / 2 - off
                for j in part:
                    acc = trunc32(acc + a_hi[:, 8 * j:8 * j + 8] @ W_hi[:, 8 * j:8 * j + 8].T)
                z = rn32(z + acc)
            if bias_out:
                z = rn32(z + bvec)
        i_, f_, g_, o_ = sigmoid(z[:, :H]), sigmoid(z[:, H:2 * H]), np.tanh(z[:, 2 * H:3 * H]), sigmoid(z[:, 3 * H:])
        i_, f_, g_, o_ = rn32(i_), rn32(f_), rn32(g_), rn32(o_)
        c = rn32(f_ * c + i_ * g_)
        h = rn32(o_ * np.tanh(c))
        Hs.append(h); Cs.append(c); Gs.append((i_, f_, g_, o_))
    return g, w, np.array(Hs), np.array(Cs), Gs, (B, T, I, H)


def bptt(g, w, Hs, Cs, Gs, dims):
    """exact fp64 BPTT on the given tapes (seq-to-one NSE loss of the golden case) -> dW_ih, dW_hh, db"""
    B, T, I, H = dims
    x = g["x_lstm"].astype(np.float64)
    lw, lb = w["linear.weight"].astype(np.float64), w["linear.bias"].astype(np.float64)
    Whh = w["lstm.weight_hh_l0"].astype(np.float64)
    y = Hs[-1] @ lw.T + lb
    yo, sd = g["y_obs"][:, 0].astype(np.float64), g["basin_std"][:, 0].astype(np.float64)
    m = ~np.isnan(yo)
    dy = np.where(m, 2 * (y - np.nan_to_num(yo)) / (sd + 0.1) ** 2 / m.sum(), 0.0)
    dh = dy @ lw
    dc = np.zeros((B, H))
    dWih = np.zeros((4 * H, I)); dWhh = np.zeros((4 * H, H)); db = np.zeros(4 * H)
    for t in range(T - 1, -1, -1):
        i_, f_, g_, o_ = Gs[t]
        tc = np.tanh(Cs[t])
        dct = dc + dh * o_ * (1 - tc * tc)
        cp = Cs[t - 1] if t > 0 else np.zeros((B, H))
        hp = Hs[t - 1] if t > 0 else np.zeros((B, H))
        dz = np.concatenate([dct * g_ * i_ * (1 - i_), dct * cp * f_ * (1 - f_), dct * i_ * (1 - g_ * g_), dh * tc * o_ * (1 - o_)], axis=1)
        dWih += dz.T @ x[:, t]; dWhh += dz.T @ hp; db += dz.sum(0)
        dh = dz @ Whh
        dc = dct * f_
    return dWih, dWhh, db


if __name__ == "__main__":
    case = sys.argv[1]
    schemes = sys.argv[2:]
    gref = None
    for sch in ["exact64"] + schemes:
        parts = sch.split("+")
        comp = 0.0
        for p in parts:
            if p.startswith("comp"):
                comp = float(p[4:])
        if sch == "exact64":
            # fp64 reference tapes: the same loop without any rounding
            import torch
            g = golden(case); w = cudalstm_weights_np(g)
            B, T, I, H = [int(v) for v in g["meta"]][:4]
            xx = g["x_lstm"].astype(np.float64)
            h = np.zeros((B, H)); c = np.zeros((B, H)); Hs, Cs, Gs = [], [], []
            Wih, Whh = w["lstm.weight_ih_l0"].astype(np.float64), w["lstm.weight_hh_l0"].astype(np.float64)
            b = (w["lstm.bias_ih_l0"].astype(np.float64) + w["lstm.bias_hh_l0"].astype(np.float64))
            for t in range(T):
                z = xx[:, t] @ Wih.T + b + h @ Whh.T
                i_, f_, g_, o_ = sigmoid(z[:, :H]), sigmoid(z[:, H:2 * H]), np.tanh(z[:, 2 * H:3 * H]), sigmoid(z[:, 3 * H:])
                c = f_ * c + i_ * g_; h = o_ * np.tanh(c)
                Hs.append(h); Cs.append(c); Gs.append((i_, f_, g_, o_))
            gref = bptt(g, w, np.array(Hs), np.array(Cs), Gs, (B, T, I, H))
            Href = np.array(Hs)
            continue
        g, w, Hs, Cs, Gs, dims = run(case, parts, comp)
        got = bptt(g, w, Hs, Cs, Gs, dims)
        rel = lambda a, b: np.max(np.abs(a - b)) / np.max(np.abs(b))
        print(f"{sch:28s} h err {rel(Hs, Href):.2e}  dW_ih {rel(got[0], gref[0]):.2e} dW_hh {rel(got[1], gref[1]):.2e} db {rel(got[2], gref[2]):.2e}", flush=True)
