"""Deterministic synthetic tree-state batches of the shapes BASELINE.json names.

This is input generation only (bench / tests); it evaluates nothing on the hot path.  It is written in
device-agnostic torch so the same code makes small CPU batches for the parity tests and full-size
batches directly in HBM for the bench.

Conventions follow the reference's simulators (ForwardTimeSimulatorResub.java:297-308,
BirthDeathStubSimulator.java:366-377): leaves are nodes 0..n-1, internal nodes follow, the root is
node N-1 (N = 2n-1); a sampled ancestor is a zero-length leaf under a "fake" internal node; stub
counts live on branches (root: none).
"""
from dataclasses import dataclass

import torch

from . import _abi as abi

SEEDS = {"C2": 20261019, "C3": 20261020, "C4": 20261021, "C5": 20261022}


@dataclass
class Batch:
    n_trees: int
    n_taxa: int
    n_nodes: int
    stride: int
    flags: int
    stub_mode: int
    height: torch.Tensor      # [T, stride] f64
    parent: torch.Tensor      # [T, stride] i32, -1 = root
    rate: torch.Tensor        # [T, stride] f64
    spike: torch.Tensor       # [T, stride] f64
    nstubs: torch.Tensor      # [T, stride] i32
    params: torch.Tensor      # [T, NPARAM] f64
    use_spike: torch.Tensor   # [T] u8
    kind: str = ""

    def to(self, device):
        kw = {k: (v.to(device) if isinstance(v, torch.Tensor) else v) for k, v in self.__dict__.items()}
        return Batch(**kw)

    def rows(self, t0, t1):
        """trees [t0, t1) as a Batch of views (no copy): a shard or a chunk of a resident batch"""
        kw = {k: (v[t0:t1] if isinstance(v, torch.Tensor) else v) for k, v in self.__dict__.items()}
        kw["n_trees"] = t1 - t0
        return Batch(**kw)

    def dense(self, name):
        """[T, n_nodes] contiguous host numpy view of a branch array (the layout the C ABI takes)."""
        return getattr(self, name)[:, : self.n_nodes].contiguous().cpu().numpy()

    def branch_count(self):
        return self.n_trees * self.n_nodes


DEFAULT_FLAGS = (abi.F_HAS_INDICATOR | abi.F_HAS_RATES | abi.F_STUBS_TO_CLOCK | abi.F_STUBS_TO_SPIKE_PRIOR
                 | abi.F_STUBS_TO_TREE_PRIOR)


def device_stride(n_nodes):
    return max(32, (n_nodes + 31) // 32 * 32)


def _topology(T, n, gen, device):
    """Random binary trees by uniformly joining pairs of active lineages (Yule-like).  Internal node
    n+k is created at step k, so node numbers increase with height and the root is node 2n-2."""
    N = 2 * n - 1
    parent = torch.full((T, N), -1, dtype=torch.int64, device=device)
    active = torch.arange(n, device=device).expand(T, n).clone()
    ar = torch.arange(T, device=device)
    u = torch.rand((n - 1, T, 2), generator=gen, device=device, dtype=torch.float64) if n > 1 else None
    for k in range(n - 1):
        m = n - k
        i = (u[k, :, 0] * m).long().clamp_(max=m - 1)
        j = (u[k, :, 1] * (m - 1)).long().clamp_(max=m - 2)
        j = j + (j >= i).long()
        lo = torch.minimum(i, j)
        hi = torch.maximum(i, j)
        a = active[ar, lo]
        b = active[ar, hi]
        new = n + k
        parent[ar, a] = new
        parent[ar, b] = new
        active[ar, lo] = new
        active[ar, hi] = active[ar, m - 1]
    return parent


def _expected_stubs(h1, h0, lam, mu, psi, rho):
    """Poisson mean of the stub count on a branch (StumpedTreePrior.java:749-769), vectorised; only
    used to draw plausible stub counts."""
    lam, mu, psi, rho = (x.unsqueeze(1) for x in (lam, mu, psi, rho))
    c1 = torch.sqrt(torch.abs((lam - mu - psi) ** 2 + 4 * lam * psi))
    c2 = -(lam - mu - 2 * lam * rho - psi) / c1
    f = ((c2 - 1) * torch.exp(-c1 * h1) - c2 - 1) / ((c2 - 1) * torch.exp(-c1 * h0) - c2 - 1)
    e_s = (h0 - h1) * (lam + mu + psi - c1) + 2 * torch.log(f)
    lb1 = torch.log(lam * torch.exp((lam - mu) * h1) - mu)
    lb0 = torch.log(lam * torch.exp((lam - mu) * h0) - mu)
    e_n = 2 * ((lb1 - lb0) + lam * (h0 - h1))
    nos = (psi == 0) & (rho == 1)
    e = torch.where(nos, e_n, e_s)
    return torch.nan_to_num(e, nan=0.0, posinf=0.0, neginf=0.0).clamp_(min=0.0, max=50.0)


def make_batch(kind, n_trees, n_taxa, seed=None, device="cpu", tree_offset=0, shared_tree=False):
    """kind: 'C2' (BDWS: psi=0, rho=1, ultrametric, spikes on), 'C3' (FBDWS: psi>0, rho<1, dated tips,
    5% sampled ancestors, indicator Bernoulli(0.5) per state), 'C4' (ultrametric; rho=1 for even trees,
    U(0.1,1) for odd trees)."""
    kind = kind.upper()
    if kind not in ("C2", "C3", "C4"):
        raise ValueError("kind must be C2, C3 or C4")
    if seed is None:
        seed = SEEDS[kind]
    device = torch.device(device)
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed) + 7919 * int(tree_offset))
    T, n = int(n_trees), int(n_taxa)
    N = 2 * n - 1
    S = device_stride(N)
    f64 = torch.float64

    def rand(*shape):
        return torch.rand(shape, generator=gen, device=device, dtype=f64)

    def randn(*shape):
        return torch.randn(shape, generator=gen, device=device, dtype=f64)

    def expo(*shape):
        return -torch.log1p(-rand(*shape))

    # ---- per-tree parameters (template defaults / priors, fxtemplates/GammaSpikeClockModel.xml:31-70) ----
    lam = torch.exp(0.5 * randn(T)).clamp_(0.05, 5.0)
    r0 = 1.0 + expo(T)
    mu = lam / r0
    if kind == "C3":
        # Beta(1,5) via inverse CDF: 1 - (1-u)^(1/5)
        s = 1.0 - torch.pow(1.0 - rand(T), 0.2)
        s = s.clamp_(1e-3, 0.9)
        psi = s * mu / (1 - s)
        rho = 0.1 + 0.9 * rand(T)
    elif kind == "C4":
        psi = torch.zeros(T, dtype=f64, device=device)
        odd = ((torch.arange(T, device=device) + tree_offset) % 2) == 1
        rho = torch.where(odd, 0.1 + 0.9 * rand(T), torch.ones(T, dtype=f64, device=device))
    else:
        psi = torch.zeros(T, dtype=f64, device=device)
        rho = torch.ones(T, dtype=f64, device=device)
    shape = torch.exp(torch.log(torch.tensor(2.0, dtype=f64)) - 0.125 + 0.5 * randn(T)).clamp_(min=0.2)
    spike_mean = torch.exp(torch.log(torch.tensor(0.01, dtype=f64)) - 0.72 + 1.2 * randn(T)).clamp_(1e-6, 1.0)
    clock_sd = torch._standard_gamma(torch.full((T,), 5.0, dtype=f64, device=device), generator=gen) * 0.05
    clock_rate = torch.ones(T, dtype=f64, device=device)

    # ---- topology + ultrametric heights ----
    parent = _topology(T, n, gen, device)
    height = torch.zeros((T, N), dtype=f64, device=device)
    if n > 1:
        lineages = torch.arange(n, 1, -1, device=device, dtype=f64)           # n, n-1, ..., 2
        inc = expo(T, n - 1) / lineages
        ih = torch.cumsum(inc, dim=1)
        t_target = torch.minimum(1.0 + 19.0 * rand(T), 50.0 / lam)
        ih = ih * (t_target / ih[:, -1]).unsqueeze(1)
        height[:, n:] = ih
    root_h = height[:, N - 1].clone()

    ar = torch.arange(T, device=device).unsqueeze(1)
    leaf_parent = parent[:, :n].clamp(min=0)
    is_sa = torch.zeros((T, N), dtype=torch.bool, device=device)
    if kind == "C3" and n > 1:
        ph = height[ar, leaf_parent]                                          # parent heights of leaves
        dated = rand(T, n) >= 0.6
        lh = torch.where(dated, 0.8 * rand(T, n) * ph, torch.zeros_like(ph))
        # sampled ancestors: ~5% of leaves, at most one per parent; height == parent height exactly
        pick = rand(T, n) < 0.05
        idx = torch.arange(n, device=device).expand(T, n)
        big = torch.full((T, N), n, dtype=torch.int64, device=device)
        first = big.scatter_reduce(1, leaf_parent, torch.where(pick, idx, torch.full_like(idx, n)), reduce="amin")
        sa = pick & (first[ar, leaf_parent] == idx)
        lh = torch.where(sa, ph, lh)
        height[:, :n] = lh
        is_sa[:, :n] = sa

    if shared_tree:   # config 5: one tree (and its tree-prior parameters) for every chain state
        parent = parent[:1].expand(T, N).contiguous()
        height = height[:1].expand(T, N).contiguous()
        is_sa = is_sa[:1].expand(T, N).contiguous()
        root_h = root_h[:1].expand(T).contiguous()
        lam, mu, psi, rho = (x[:1].expand(T).contiguous() for x in (lam, mu, psi, rho))

    # ---- derived predicates (as Node.isDirectAncestor / isFake define them) ----
    par_c = parent.clamp(min=0)
    hp = height[ar, par_c]
    is_root = parent < 0
    hp = torch.where(is_root, height, hp)
    fake = torch.zeros((T, N), dtype=torch.int64, device=device).scatter_reduce(
        1, par_c, (is_sa & ~is_root).long(), reduce="amax").bool()      # a fake node has a direct-ancestor child
    parent_fake = fake[ar, par_c] & ~is_root

    # ---- per-branch values ----
    e = _expected_stubs(height, hp, lam, mu, psi, rho)
    e = torch.where(is_root | is_sa | (hp <= height), torch.zeros_like(e), e)
    nstubs = torch.poisson(e, generator=gen).to(torch.int64)
    nstubs[:, N - 1] = 0
    m = torch.where(is_root, nstubs + 1, torch.where(is_sa, torch.zeros_like(nstubs),
                                                      torch.where(parent_fake, nstubs, nstubs + 1)))
    alpha = (shape.unsqueeze(1) * m.to(f64)).clamp_(min=1e-3)
    g = torch._standard_gamma(alpha, generator=gen) / shape.unsqueeze(1)
    spike = torch.where(m > 0, g.clamp_(min=1e-300), torch.zeros_like(g))
    sd = clock_sd.unsqueeze(1)
    rate = torch.exp(-0.5 * sd * sd + sd * randn(T, N))

    if kind == "C3":
        use_spike = (rand(T) < 0.5).to(torch.uint8)
    else:
        use_spike = torch.ones(T, dtype=torch.uint8, device=device)

    params = torch.stack([clock_rate, spike_mean, shape, clock_sd, lam, mu, psi, rho, root_h * 1.1], dim=1).contiguous()

    def pad(x, dtype, fill=0):
        out = torch.full((T, S), fill, dtype=dtype, device=device)
        out[:, :N] = x.to(dtype)
        return out

    return Batch(n_trees=T, n_taxa=n, n_nodes=N, stride=S, flags=DEFAULT_FLAGS, stub_mode=abi.STUBS_COUNTS,
                 height=pad(height, f64), parent=pad(parent, torch.int32, -1), rate=pad(rate, f64, 1.0),
                 spike=pad(spike, f64), nstubs=pad(nstubs, torch.int32), params=params, use_spike=use_spike, kind=kind)


def make_batch_chunked(kind, n_trees, n_taxa, seed=None, device="cpu", chunk=16384):
    """Same distribution as make_batch, generated in chunks to bound temporary memory (full-size bench batches): the arrays
    of the whole batch are allocated once and filled chunk by chunk, so the peak is the batch plus one chunk (the 1M-tree
    configuration is 128 GB of inputs on one 180 GB GPU)."""
    if n_trees <= chunk:
        return make_batch(kind, n_trees, n_taxa, seed=seed, device=device)
    out = None
    for t0 in range(0, n_trees, chunk):
        part = make_batch(kind, min(chunk, n_trees - t0), n_taxa, seed=seed, device=device, tree_offset=t0)
        if out is None:
            alloc = lambda x: torch.empty((n_trees,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
            out = Batch(n_trees=n_trees, n_taxa=part.n_taxa, n_nodes=part.n_nodes, stride=part.stride, flags=part.flags,
                        stub_mode=part.stub_mode, height=alloc(part.height), parent=alloc(part.parent), rate=alloc(part.rate),
                        spike=alloc(part.spike), nstubs=alloc(part.nstubs), params=alloc(part.params),
                        use_spike=alloc(part.use_spike), kind=part.kind)
        t1 = t0 + part.n_trees
        for name in ("height", "parent", "rate", "spike", "nstubs", "params", "use_spike"):
            getattr(out, name)[t0:t1] = getattr(part, name)
        del part
    return out


def move_root(batch, j):
    """In place: node labels n-1 (the root of the generated trees) and j (an internal node) exchanged in every tree.  BEAST's
    tree operators leave the root at any internal node number; the generators here (like the reference's simulators) put it
    last."""
    n1 = batch.n_nodes - 1
    if not (batch.n_taxa <= j < n1):
        raise ValueError("j must be an internal node other than the last")
    for name in ("height", "rate", "spike", "nstubs", "parent"):
        x = getattr(batch, name)
        t = x[:, n1].clone()
        x[:, n1] = x[:, j]
        x[:, j] = t
    p = batch.parent
    is1, isj = p == n1, p == j
    p[is1] = j
    p[isj] = n1
    return batch


def make_chain_states(n_chains, n_taxa, seed=None, device="cpu", kind="C2"):
    """BASELINE config 5: ONE tree (topology, heights, tree-prior parameters) replicated to `n_chains` chain states; every
    chain has its own branch rates, spikes, stub counts (drawn for that tree) and clock / spike hyper-parameters
    ("one 500-taxon tree replicated to 64 chain states")."""
    b = make_batch(kind, n_chains, n_taxa, seed=SEEDS["C5"] if seed is None else seed, device=device, shared_tree=True)
    b.kind = "C5"
    return b


def make_proposals(a, n_props, seed=None, hyper=True):
    """BASELINE config 5: a batch of MCMC proposals on the host arrays `a` (dense [T][N] numpy arrays as the C ABI takes
    them), drawn from the operator mix of the reference's template (GammaSpikeClockModel.xml:80-112):
    ~80 % touch <= 4 branches, ~20 % touch every branch or a per-tree scalar.

        35 %  SpikeUpRateDown (SpikeUpRateDown.java:23-47): one branch, rate * s, spike / s
        15 %  SpikeFlipOperator (SpikeFlipOperator.java:30-67): one spike redrawn
        20 %  StumpedTreeConstantDistanceOperator (:61-183): one internal height + the three adjacent rates
        10 %  stub count +-1 on one branch
        10 %  StumpedTreeScaler (:63-161): Tree.scale(s) (every internal node that is not fake) through height_scale,
              with lambda scaled down as the template's down-parameter
        10 %  hyper-parameter / indicator move (scale operators on GSMspikeShape, GSMclockSD, lambda; BitFlip on
              GSMuseSpikeModel): per-tree scalars replaced (only when hyper=True)

    Returns the gsm_eval_dirty arguments as a dict of numpy arrays (CSR)."""
    import numpy as np
    rng = np.random.default_rng(SEEDS["C5"] if seed is None else seed)
    height, parent = a["height"], a["parent"]
    T, N = height.shape
    L = (N + 1) // 2
    tree_of_prop = rng.integers(0, T, n_props).astype(np.int32)
    kind = rng.choice(6 if hyper else 5, n_props, p=[.35, .15, .20, .10, .10, .10] if hyper else [.35 / .9, .15 / .9, .20 / .9, .10 / .9, .10 / .9])
    ptr = np.zeros(n_props + 1, np.int64)
    nodes, nh, nr, ns, nk = [], [], [], [], []
    new_params = a["params"][tree_of_prop].copy()
    new_use = a["use_spike"][tree_of_prop].copy()
    scale = np.ones(n_props, np.float64)
    for p in range(n_props):
        t = int(tree_of_prop[p])
        h, par, rate, spike, nst = height[t], parent[t], a["rate"][t], a["spike"][t], a["nstubs"][t]

        def put(i, hh=None, rr=None, ss=None, kk=None):
            nodes.append(i)
            nh.append(h[i] if hh is None else hh)
            nr.append(rate[i] if rr is None else rr)
            ns.append(spike[i] if ss is None else ss)
            nk.append(nst[i] if kk is None else kk)
        k = int(kind[p])
        if k == 0:
            i = int(rng.integers(0, N - 1))
            s = float(np.exp(0.3 * rng.standard_normal()))
            put(i, rr=rate[i] * s, ss=spike[i] / s)
        elif k == 1:
            i = int(rng.integers(0, N - 1))
            put(i, ss=(0.0 if spike[i] == 0.0 else float(rng.exponential(0.5))))
        elif k == 2:
            i = int(rng.integers(L, N))
            ch = np.nonzero(par == i)[0]
            lo = float(h[ch].max()) if len(ch) else 0.0
            hi = float(h[par[i]]) if par[i] >= 0 else float(h[i]) * 1.2 + 1e-3
            put(i, hh=lo + (hi - lo) * float(rng.uniform(0.05, 0.95)), rr=rate[i] * float(rng.uniform(0.8, 1.25)))
            for c in ch[:2]:
                put(int(c), rr=rate[c] * float(rng.uniform(0.8, 1.25)))
        elif k == 3:
            i = int(rng.integers(0, N - 1))
            put(i, kk=max(0, int(nst[i]) + (1 if rng.random() < 0.5 else -1)))
        elif k == 4:
            scale[p] = float(rng.uniform(0.9, 1.1))
            if hyper:
                new_params[p, abi.P_LAMBDA] /= scale[p]
        else:
            which = int(rng.integers(0, 4))
            if which == 0:
                new_params[p, abi.P_SPIKE_SHAPE] *= float(rng.uniform(0.8, 1.25))
            elif which == 1:
                new_params[p, abi.P_CLOCK_SD] *= float(rng.uniform(0.8, 1.25))
            elif which == 2:
                new_params[p, abi.P_LAMBDA] *= float(rng.uniform(1.0, 1.25))
            else:
                new_use[p] = 1 - new_use[p]
        ptr[p + 1] = len(nodes)
    out = dict(tree_of_prop=tree_of_prop, dirty_ptr=ptr, dirty_nodes=np.asarray(nodes, np.int32),
               new_height=np.asarray(nh, np.float64), new_rate=np.asarray(nr, np.float64),
               new_spike=np.asarray(ns, np.float64), new_nstubs=np.asarray(nk, np.int32), height_scale=scale)
    if hyper:
        out["new_params"] = new_params
        out["new_use_spike"] = new_use.astype(np.uint8)
    return out
