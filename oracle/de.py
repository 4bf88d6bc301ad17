"""
Oracle of the differential-evolution generation logic.  TEST INFRASTRUCTURE ONLY.

The reference calls `scipy.optimize.differential_evolution` (calibrate/calibrate_abcd.py:103-110: best1bin,
Latin-hypercube init, dither (0.5, 1), recombination 0.7, tol 0.01, polish=False).  scipy is a third-party
dependency that is not part of /root/reference; the version installed here (and on the GPU box) is 1.18.1.
This file

  1. restates, in numpy, the generation logic the CUDA kernels of xanthos_b200/csrc/de.cu implement, draw by draw
     (Philox4x32-10 keyed by the seed, counter = (generation, problem, member-or-gene, stream));
  2. PINS that restatement to scipy itself: `ScipyReplay` drives scipy's own, unmodified
     `DifferentialEvolutionSolver` (strategy 'best1bin', updating='deferred') with a random-number object that
     replays the Philox stream in the order scipy consumes it, and checks that the trial vectors, the selection
     and the convergence decision of every generation are BITWISE those of (1).  tests/test_oracle.py runs it on
     the CPU; the GPU tests then compare the kernels with (1) bitwise.

What cannot be bitwise: scipy's default updating='immediate' consumes trial energies member by member; the
device solver (like scipy's own updating='deferred') evaluates a generation as one batch.  The pin is therefore
on scipy's deferred mode.  The Latin hypercube of `lhs_init` uses its own draw order (scipy's
`init='latinhypercube'` draws from the same distribution); the replay starts from the same initial population.

Scaling: scipy maps x in [0, 1] to parameters as 0.5 (lo + hi) + (x - 0.5) |hi - lo|, the kernels as
lo + x (hi - lo).  For the bounds (-0.5, 0.5) both are fl(x - 0.5): the replay uses those bounds, so that the
objective sees bit-identical parameter vectors on both sides.
"""

import sys

import numpy as np

STREAM_INIT, STREAM_PICK, STREAM_CROSS, STREAM_OOB, STREAM_DITHER = 1, 2, 3, 4, 5
DE_MAX_S, DE_MAX_D = 256, 8
_M0, _M1, _W0, _W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 (Salmon et al. 2011) on arrays of uint32 counters; returns four uint32 arrays."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint64) & 0xffffffff for c in np.broadcast_arrays(c0, c1, c2, c3)]
    a, b = np.uint64(k0 & 0xffffffff), np.uint64(k1 & 0xffffffff)
    m32 = np.uint64(0xffffffff)
    for _ in range(10):
        p0 = np.uint64(_M0) * c0
        p1 = np.uint64(_M1) * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & m32
        hi1, lo1 = p1 >> np.uint64(32), p1 & m32
        c0, c1, c2, c3 = hi1 ^ c1 ^ a, lo1, hi0 ^ c3 ^ b, lo0
        a = (a + np.uint64(_W0)) & m32
        b = (b + np.uint64(_W1)) & m32
    return c0, c1, c2, c3


def u01(hi, lo):
    """53-bit uniform in [0, 1) from two uint32 words (de.cu:u01)."""
    return ((((hi << np.uint64(32)) | lo) >> np.uint64(11)).astype(np.float64)) * (1.0 / 9007199254740992.0)


def split_seed(seed):
    seed = int(seed)
    return seed & 0xffffffff, (seed >> 32) & 0xffffffff


def lhs_init(n_problems, S, D, seed):
    """de_init_kernel: per (problem, dimension) a Fisher-Yates order of the S strata, one uniform per stratum."""
    k0, k1 = split_seed(seed)
    pop = np.empty((n_problems, S, D))
    seg = 1.0 / S
    for prob in range(n_problems):
        for dim in range(D):
            perm = np.arange(S)
            ks = np.arange(S - 1, 0, -1)
            r = philox4x32_10(0, prob, dim * DE_MAX_S + ks, STREAM_INIT, k0, k1)
            js = np.floor(u01(r[0], r[1]) * (ks + 1)).astype(int)
            for k, j in zip(ks, js):
                perm[k], perm[j] = perm[j], perm[k]
            r = philox4x32_10(1, prob, dim * DE_MAX_S + np.arange(S), STREAM_INIT, k0, k1)
            pop[prob, :, dim] = seg * u01(r[0], r[1]) + seg * perm
    return pop


def draws(prob, gen, S, D, seed, mutation=(0.5, 1.0)):
    """Every random quantity de_trial_kernel uses for (problem, generation)."""
    k0, k1 = split_seed(seed)
    i = np.arange(S)
    rd = philox4x32_10(gen, prob, 0, STREAM_DITHER, k0, k1)
    scale = mutation[0] + (mutation[1] - mutation[0]) * float(u01(rd[0], rd[1]))
    rp = philox4x32_10(gen, prob, i, STREAM_PICK, k0, k1)
    r0 = (i + 1 + np.floor(u01(rp[0], rp[1]) * (S - 1)).astype(int)) % S
    r1 = (i + 1 + np.floor(u01(rp[2], rp[3]) * (S - 2)).astype(int)) % S
    r1 = np.where(r1 == r0, (r1 + 1) % S, r1)
    r1 = np.where(r1 == i, (r1 + 1) % S, r1)
    r1 = np.where(r1 == r0, (r1 + 1) % S, r1)
    words = []
    for blk in range(3):
        words += list(philox4x32_10(gen, prob, i + blk * DE_MAX_S, STREAM_CROSS, k0, k1))
    bits = np.stack(words, axis=1)                                              # [S, 12] uint64 holding uint32 words
    forced = np.floor(u01(bits[:, 10], bits[:, 11]) * D).astype(int)
    cross_u = bits[:, :D].astype(np.float64) * (1.0 / 4294967296.0)             # [S, D]
    jj = np.arange(D)
    ro = philox4x32_10(gen, prob, i[:, None] * DE_MAX_D + jj[None, :], STREAM_OOB, k0, k1)
    oob_u = u01(ro[0], ro[1])                                                   # [S, D]
    return dict(scale=scale, r0=r0, r1=r1, forced=forced, cross_u=cross_u, oob_u=oob_u)


def trial(pop, energy, prob, gen, seed, mutation=(0.5, 1.0), recombination=0.7):
    """de_trial_kernel for one problem: pop [S, D] in [0, 1], energy [S] -> trial vectors [S, D]."""
    S, D = pop.shape
    d = draws(prob, gen, S, D, seed, mutation)
    e = np.where(np.isnan(energy), np.inf, energy)
    best = int(np.argmin(e))
    mutant = pop[best][None, :] + d['scale'] * (pop[d['r0']] - pop[d['r1']])
    cross = d['cross_u'] < recombination
    cross[np.arange(S), d['forced']] = True
    x = np.where(cross, mutant, pop)
    oob = (x < 0.0) | (x > 1.0)
    return np.where(oob, d['oob_u'], x), d


def select(pop, energy, trial_x, trial_e):
    """de_select_kernel: deferred selection, trial kept where its energy <= the member's (NaN = +inf)."""
    e = np.where(np.isnan(energy), np.inf, energy)
    et = np.where(np.isnan(trial_e), np.inf, trial_e)
    better = et <= e
    return np.where(better[:, None], trial_x, pop), np.where(better, et, e)


def converged(energy, tol=0.01, atol=0.0):
    """scipy's DifferentialEvolutionSolver.converged()."""
    if np.any(np.isinf(energy)):
        return False
    return bool(np.std(energy) <= atol + tol * np.abs(np.mean(energy)))


class _ReplayRNG:
    """
    Stands in for the numpy Generator inside scipy's solver and hands it the Philox draws of the generation that is
    being replayed, in scipy's own consumption order (scipy 1.18.1, `__next__` / `_mutate_many` / `_select_samples` /
    `_ensure_constraint`, deferred updating):
        uniform(lo, hi)             dither (one per generation)
        shuffle(index array)        once per candidate: the first two entries become the two picked members
        integers(0, D, size=S)      the forced gene of every candidate
        uniform(size=(S, D))        crossover uniforms
        uniform(size=n_oob)         replacement of out-of-bounds genes, in C order of the mask
    scipy keeps its best member at index 0 by swapping rows; `order[k]` is the kernel-side member held in scipy's
    row k, so that every draw is delivered to the row that holds the member it belongs to.
    """

    def __init__(self):
        self.d = None
        self.order = None
        self.cand = 0

    def begin(self, d, order):
        self.d, self.order, self.cand = d, np.asarray(order), 0
        self.inv = np.argsort(self.order)

    def uniform(self, low=0.0, high=1.0, size=None):
        if size is None:
            return self.d['scale']                                             # == low + (high - low) * u, bit for bit
        if isinstance(size, tuple):
            return self.d['cross_u'][self.order]
        mask = sys._getframe(1).f_locals['mask']                               # _ensure_constraint's local
        assert int(np.count_nonzero(mask)) == int(size)
        return self.d['oob_u'][self.order][mask]

    def shuffle(self, arr):
        member = self.order[self.cand]
        a, b = self.inv[self.d['r0'][member]], self.inv[self.d['r1'][member]]
        rest = [k for k in range(len(arr)) if k != a and k != b]
        arr[:] = [a, b] + rest
        self.cand += 1

    def integers(self, low, high=None, size=None, dtype=np.int64, endpoint=False):
        return self.d['forced'][self.order]

    randint = integers                          # scipy's rng_integers falls back to the RandomState spelling


class ScipyReplay:
    """
    scipy's own DifferentialEvolutionSolver (deferred updating) on `func`, replaying the Philox stream of
    (seed, problem).  `step()` advances one generation and returns scipy's population / energies re-ordered to the
    kernel-side member order, plus the parameter vectors scipy evaluated (in member order).
    """

    def __init__(self, func, pop0, prob, seed, mutation=(0.5, 1.0), recombination=0.7, tol=0.01, atol=0.0):
        from scipy.optimize._differentialevolution import DifferentialEvolutionSolver
        S, D = pop0.shape
        self.S, self.D, self.prob, self.seed, self.mutation = S, D, prob, seed, mutation
        self.calls = []

        def recorded(x):
            self.calls.append(np.array(x, dtype=float))
            return func(x)
        self.rng = _ReplayRNG()
        self.solver = DifferentialEvolutionSolver(recorded, [(-0.5, 0.5)] * D, strategy='best1bin', maxiter=10 ** 6,
                                                  popsize=max(1, S // D), tol=tol, atol=atol, mutation=mutation,
                                                  recombination=recombination, polish=False,
                                                  init=np.array(pop0) - 0.5, updating='deferred')
        assert self.solver.num_population_members == S
        self.solver.random_number_generator = self.rng
        self.solver.population[:] = pop0                                       # exact (init goes through an unscale)
        self.order = np.arange(S)
        # initial energies, as solve() / __next__ compute them on first use
        self.solver.population_energies[:] = [recorded(x - 0.5) for x in pop0]
        self.solver._nfev = S
        self.calls = []
        self._promote()
        self.gen = 0

    def _promote(self):
        before = self.solver.population.copy()
        self.solver._promote_lowest_energy()
        after = self.solver.population
        if not np.array_equal(before, after):
            ch = np.nonzero((before != after).any(axis=1))[0]
            assert len(ch) == 2
            self.order[ch] = self.order[ch[::-1]]

    def state(self):
        inv = np.argsort(self.order)
        return self.solver.population[inv].copy(), self.solver.population_energies[inv].copy()

    def step(self):
        """One generation of scipy's solver; returns the parameter vectors it evaluated, in member order."""
        self.gen += 1
        d = draws(self.prob, self.gen, self.S, self.D, self.seed, self.mutation)
        self.rng.begin(d, self.order)
        self.calls = []
        order_before = self.order.copy()
        next(self.solver)                       # trial vectors, evaluation, selection, best member to row 0
        evaluated = np.stack(self.calls)        # in scipy's row order at the time of the evaluation
        return evaluated[np.argsort(order_before)]

    def resync(self, pop_member_order):
        """After a step: derive which member each scipy row holds by matching rows with the kernel-side population."""
        rows = self.solver.population
        order = np.full(self.S, -1)
        taken = np.zeros(self.S, dtype=bool)
        for k in range(self.S):
            hit = np.nonzero((pop_member_order == rows[k]).all(axis=1) & ~taken)[0]
            assert len(hit) >= 1, "scipy row {} is not a member of the restated population".format(k)
            order[k] = hit[0]
            taken[hit[0]] = True
        self.order = order


def replay_against_scipy(func, n_gen, S=25, D=5, prob=0, seed=12345, tol=0.01):
    """
    Run the restated solver and scipy's solver side by side for `n_gen` generations; returns a list of per-generation
    dicts with the bitwise comparisons (all must be True) and the number of out-of-bounds genes that were redrawn.
    """
    pop = lhs_init(prob + 1, S, D, seed)[prob]
    energy = np.array([func(x - 0.5) for x in pop])
    rep = ScipyReplay(func, pop, prob, seed, tol=tol)
    out = []
    for gen in range(1, n_gen + 1):
        tx, d = trial(pop, energy, prob, gen, seed)
        te = np.array([func(x - 0.5) for x in tx])
        evaluated = rep.step()
        pop, energy = select(pop, energy, tx, te)
        rep.resync(pop)
        sp, se = rep.state()
        mutant_oob = int(np.count_nonzero(tx == d['oob_u']))
        out.append(dict(gen=gen, trial_equal=bool(np.array_equal(evaluated, tx - 0.5)),
                        pop_equal=bool(np.array_equal(sp, pop)), energy_equal=bool(np.array_equal(se, energy)),
                        converged_equal=rep.solver.converged() == converged(energy, tol), n_oob=mutant_oob,
                        best_at_row0=bool(rep.order[0] == int(np.argmin(energy)))))
    return out
