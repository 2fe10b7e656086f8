"""CPU restatement of HOW k3_filtered_kernel (sdvar_b200/csrc/sampling.cu, v3) locates the top-k / top-p cuts -- value-bin
histograms of (count, fixed-point E), one scan from the top, the bins a top-p crossing can fall into for any exact top-k cut,
all-pairs ranks among the few candidates, and the "cut below the candidate range" rule -- checked against the DEFINITION of
the cuts (oracle/spec_c: sdvar_spec_sample) on random, peaked, tied and quantised rows.  The kernel's bit-exactness itself is
tested on the GPU (tests/test_kernels_gpu.py); this test pins the algorithm's logic, which no input may break, where there is
no GPU.  TEST INFRASTRUCTURE (imports oracle/)."""
import numpy as np
import pytest
import torch

NB, CAND_MAX = 1024, 128


def fkey(x):
    b = (np.asarray(x, np.float32) + np.float32(0.0)).view(np.uint32).astype(np.int64)
    return np.where(b & 0x80000000, (~b) & 0xFFFFFFFF, b | 0x80000000)


def thr_mass(Zi, thr_fix):
    return (Zi * thr_fix) >> 30


def fixed_E(x):
    from oracle import spec
    m = x.max()
    e = spec.expf(torch.from_numpy(x - m)).numpy()
    return [int(round(float(v) * 2.0 ** 40)) for v in e]          # exact product, round-half-even == llrintf


def cuts_by_definition(x, E, top_k, thr_fix):
    """kept mask and Z2i straight from the spec's definition"""
    V = len(x)
    keys = fkey(x)
    alive = np.ones(V, bool)
    if 0 < top_k < V:
        K = np.sort(keys)[V - top_k]
        alive = keys >= K
    Zi = sum(E[v] for v in range(V) if alive[v])
    kept = alive.copy()
    if thr_fix is not None:
        thrE = thr_mass(Zi, thr_fix)
        kmax = keys.max()
        T = 0
        for k in np.unique(keys[alive]):
            if sum(E[v] for v in range(V) if alive[v] and keys[v] <= k) <= thrE:
                T = k
            else:
                break
        kept = alive & ~((keys <= T) & (keys != kmax))
    return kept, sum(E[v] for v in range(V) if kept[v])


def cuts_like_the_kernel(x, E, top_k, thr_fix):
    """returns (klow key, Z2i) or None where the kernel takes its generic path"""
    V = len(x)
    use_k, use_p = 0 < top_k < V, thr_fix is not None
    m, xmin = x.max(), x.min()
    rng = np.float32(m) - np.float32(xmin)
    if not (rng > 0 and np.isfinite(rng)) or (use_p and thr_fix >= 1 << 30):
        return None
    bins = np.floor((x.astype(np.float64) - float(xmin)) * ((NB - 2) / float(rng))).astype(int)      # any monotone binning will do
    assert bins.min() >= 0 and bins.max() < NB
    cnt = np.bincount(bins, minlength=NB)
    if cnt.max() > 4095:
        return None
    Eb = [0] * NB
    for v in range(V):
        Eb[bins[v]] += E[v]
    cab, dex = [0] * NB, [0] * (NB + 1)          # entries / E in the bins above b; dex[b] here = kernel's pdex[b + 1], din(b) = dex[b] + Eb[b]
    c = e = 0
    for b in range(NB - 1, -1, -1):
        cab[b], dex[b] = c, e
        c += cnt[b]
        e += Eb[b]
    Etot = e
    din = lambda b: dex[b] + Eb[b]
    bK = -2
    if use_k:
        bK = next(b for b in range(NB - 1, -1, -1) if cab[b] < top_k <= cab[b] + cnt[b])
    plo = phi = NB
    if use_p:
        zmin, zmax = (dex[bK], din(bK)) if use_k else (Etot, Etot)
        tmin, tmax = zmin - thr_mass(zmin, thr_fix), zmax - thr_mass(zmax, thr_fix)
        pb = [b for b in range(NB) if Eb[b] != 0 and dex[b] < tmax and din(b) >= tmin]
        plo, phi = (min(pb), max(pb)) if pb else (0x7FFFFFFF, -1)
    kgroup = use_k and bK < plo
    cand = [(v, 1 if plo <= bins[v] <= phi else 0) for v in range(V) if (kgroup and bins[v] == bK) or plo <= bins[v] <= phi]
    if len(cand) > CAND_MAX:
        return None
    keys = fkey(x)
    xK, Zi = xmin, Etot
    info = []
    for v, g in cand:
        same = [(w, E[w]) for w, gw in cand if gw == g]
        cge = sum(1 for w, _ in same if x[w] >= x[v])
        cgt = sum(1 for w, _ in same if x[w] > x[v])
        Ege = sum(Ew for w, Ew in same if x[w] >= x[v])
        Egt = Ege - (cge - cgt) * E[v]
        assert Egt == sum(Ew for w, Ew in same if x[w] > x[v])                # ties share x, hence E
        baseE = dex[phi] if g else dex[bK]
        if use_k and (g == 0 or not kgroup):
            base = cab[phi] if g else cab[bK]
            if base + cgt < top_k <= base + cge:
                xK, Zi = x[v], baseE + Ege
        info.append((v, g, Egt, baseE))
    klow, Z2i = int(fkey(xK)), Zi
    if use_p:
        target = Zi - thr_mass(Zi, thr_fix)
        tk, minkey, z2 = 0, 0xFFFFFFFF, None
        for v, g, Egt, baseE in info:
            if g != 1:
                continue
            minkey = min(minkey, int(keys[v]))
            if x[v] >= xK and x[v] < m and baseE + Egt >= target and int(keys[v]) >= tk:
                tk, z2 = int(keys[v]), baseE + Egt
        if tk:
            klow, Z2i = tk + 1, z2
        elif (kgroup if use_k else (plo > 0 and cab[plo - 1] < V)):
            klow, Z2i = minkey, dex[plo] + Eb[plo]
    return klow, Z2i


def _rows():
    g = np.random.default_rng(7)
    V = 4096
    for i in range(4):
        yield f"gauss{i}", (g.standard_normal(V) * (0.05 if i % 2 else 3.0)).astype(np.float32)
    for i in range(3):      # peaked: real checkpoints put most of the mass on a few entries
        x = (g.standard_normal(V) * 2).astype(np.float32)
        x[g.integers(0, V, 3)] += np.float32(6 + 6 * i)
        yield f"peaked{i}", x
    for i in range(3):      # quantised: tie groups around both cuts
        yield f"ties{i}", (np.round(g.standard_normal(V) * (2 + 2 * i)) / np.float32(4)).astype(np.float32)
    x = g.standard_normal(V).astype(np.float32); x[:3000] = x[3000]
    yield "one_giant_tie_group", x
    x = np.full(V, 0.25, np.float32); x[7] = 0.5
    yield "two_values", x
    x = (g.standard_normal(V) * 30).astype(np.float32)
    yield "underflowing_tail", x


@pytest.mark.parametrize("top_k,top_p", [(900, 0.96), (0, 0.96), (900, 0.0), (3000, 0.5), (1, 0.9), (4095, 0.999), (50, 0.2), (900, 0.9999999)])
def test_kernel_cut_location_equals_the_definition(top_k, top_p):
    from oracle import spec
    thr = spec.top_p_threshold(top_p)
    thr_fix = int(np.float32(thr) * np.float32(2.0 ** 30)) if thr >= 0 else None
    fast = 0
    for name, x in _rows():
        E = fixed_E(x)
        kept, Z2 = cuts_by_definition(x, E, top_k, thr_fix)
        r = cuts_like_the_kernel(x, E, top_k, thr_fix)
        if r is None:
            continue
        fast += 1
        klow, Z2i = r
        assert np.array_equal(fkey(x) >= klow, kept), (name, top_k, top_p)
        assert Z2i == Z2, (name, top_k, top_p)
    assert fast >= 8


def test_definition_restatement_equals_the_c_spec():
    """the python restatement of the definition used above keeps exactly the entries the C spec keeps"""
    from oracle import spec
    for name, x in list(_rows())[::3]:
        lg = torch.from_numpy(np.stack([x, np.zeros_like(x)])).view(2, 1, -1)
        _, mixed, _ = spec.sample(lg, [0, 1], np.float32([1.0]), np.float32([0.0]), 900, 0.96, None)
        kept, _ = cuts_by_definition(x, fixed_E(x), 900, int(np.float32(spec.top_p_threshold(0.96)) * np.float32(2.0 ** 30)))
        assert np.array_equal(torch.isfinite(mixed.view(-1)).numpy(), kept), name
