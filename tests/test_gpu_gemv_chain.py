"""GPU parity for the persistent decode-GEMV chain (csrc/gemv_chain.cu, mxq_gemv_chain_*): one launch for a
list of batch-1 gemv_mxq_forward_cuda calls (gemv_mxq_cuda.cu:225-273).  Every job is checked against the
fp64 oracle decode with north_star's per-element bound |err| <= 1e-3 * sum|w||x| (+ the fp16 store), and
against exact arithmetic on the block-floating activations the kernel really multiplies."""
import numpy as np
import pytest
import torch

from oracle import mxq_oracle as O
from tests.gpu_util import packed_to_dev
from tests.test_gpu_packed import _block_float, _outlier_x, _per_element_bound

pytestmark = pytest.mark.gpu


def _check(y, x16, p, what=""):
    ref = O.gemm_mxq_f32(x16, p)[0]
    per = _per_element_bound(x16, p)[0]
    y = y.astype(np.float64)
    bound = 1e-3 * per + np.abs(ref) * 2.0 ** -11 + 2.0 ** -25      # + half an ulp of the fp16 store (subnormal floor)
    err = np.abs(y - ref)
    assert (err <= bound).all(), f"{what}: worst err/bound {float((err / bound).max()):.3f}"
    ref_bf = (_block_float(x16) @ O.decode_mxq(p).astype(np.float64).T)[0]
    assert (np.abs(y - ref_bf) <= np.abs(ref_bf) * 2.0 ** -11 + 2.0 ** -25 + 1e-6 * per).all(), \
        f"{what}: kernel != exact integer arithmetic on the converted activations"


def _mk(cuda, shapes, seed=0):
    ps = [O.random_packed(oc, ic, seed=seed + 7 * i + oc + ic) for i, (oc, ic) in enumerate(shapes)]
    return ps, [packed_to_dev(p, cuda) for p in ps]


def test_chain_independent_mixed_shapes(cuda):
    """Ragged everything: partial 16-row tiles, a short last CTA, 1 / 2 / 3 metadata chunks per row
    (IC = 256, 4096, 8192, 11008 = 43 quad-blocks), 70B-wide rows (28672 columns: 7 chunks per tile, a
    3-stage ring), q/k/v-style jobs sharing one activation."""
    from mxq_b200 import ops
    shapes = [(256, 4096), (256, 4096), (256, 4096), (4128, 256), (128, 11008), (32, 256), (1184, 4096), (96, 8192),
              (4096, 4096), (160, 1024), (11008, 4096), (4096, 11008), (64, 28672), (32, 16384)]
    ps, pd = _mk(cuda, shapes)
    xs = {ic: _outlier_x(1, ic, seed=ic) for ic in {s[1] for s in shapes}}
    xd = {ic: torch.from_numpy(x).to(cuda) for ic, x in xs.items()}
    ys = [torch.full((1, oc), float("nan"), dtype=torch.float16, device=cuda) for oc, _ in shapes]
    chain = ops.GemvChain([(xd[ic], p, y, -1) for (oc, ic), p, y in zip(shapes, pd, ys)])
    chain.run()
    torch.cuda.synchronize()
    for i, ((oc, ic), p, y) in enumerate(zip(shapes, ps, ys)):
        _check(y.cpu().numpy()[0], xs[ic], p, f"job {i} {oc}x{ic}")
    # the kernel re-arms its own counters
    assert all(int(s.abs().sum()) == 0 for _, _, s in chain._launches)


def test_chain_reruns_and_graph_replay_are_bit_identical(cuda):
    from mxq_b200 import ops
    shapes = [(512, 4096), (1024, 4096), (512, 11008), (4096, 4096)]
    ps, pd = _mk(cuda, shapes, seed=5)
    xd = {ic: torch.from_numpy(_outlier_x(1, ic, seed=ic + 1)).to(cuda) for ic in (4096, 11008)}
    ys = [torch.zeros(oc, dtype=torch.float16, device=cuda) for oc, _ in shapes]
    chain = ops.GemvChain([(xd[ic][0], p, y, -1) for (oc, ic), p, y in zip(shapes, pd, ys)])
    chain.run()
    torch.cuda.synchronize()
    first = [y.clone() for y in ys]
    for y in ys:
        y.zero_()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        chain.run()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            chain.run()
            chain.run()
    for _ in range(3):
        for y in ys:
            y.zero_()
        g.replay()
        torch.cuda.synchronize()
        for a, b in zip(first, ys):
            assert torch.equal(a.view(torch.int16), b.view(torch.int16))


def test_chain_dependent_jobs(cuda):
    """x of job j+1 IS y of job j (ping-pong buffers are reused, so a missing ordering would read a
    half-written or stale vector): every link is checked against the oracle applied to the chain's own
    intermediate result."""
    from mxq_b200 import ops
    dims = [4096, 11008, 4096, 4096, 256, 4096]
    shapes = [(dims[i + 1], dims[i]) for i in range(len(dims) - 1)]
    ps, pd = _mk(cuda, shapes, seed=11)
    x0 = _outlier_x(1, dims[0], seed=2)
    bufs = [torch.from_numpy(x0).to(cuda)[0].clone()] + [torch.zeros(d, dtype=torch.float16, device=cuda) for d in dims[1:]]
    # scale the weights down so that five links stay inside fp16
    for p, q in zip(ps, pd):
        for k in ("scales_2nd", "scales_4b"):
            p[k] = (p[k].astype(np.float32) * 0.2).astype(np.float16)
            q[k].copy_(torch.from_numpy(p[k]))
    chain = ops.GemvChain([(bufs[i], pd[i], bufs[i + 1], i - 1) for i in range(len(shapes))])
    for rep in range(3):
        for b in bufs[1:]:
            b.fill_(float("nan"))
        chain.run()
        torch.cuda.synchronize()
        vals = [b.cpu().numpy() for b in bufs]
        assert all(np.isfinite(v).all() for v in vals)
        for i, p in enumerate(ps):
            _check(vals[i + 1], vals[i][None, :], p, f"rep {rep} link {i}")


def test_chain_extreme_activations(cuda):
    """The block-floating conversion at its edges: an all-zero vector, fp16-max groups next to zero and
    subnormal-only groups, one non-zero element per group, negative zeros."""
    from mxq_b200 import ops
    ps, pd = _mk(cuda, [(256, 4096)], seed=21)
    rng = np.random.default_rng(5)
    xs = []
    x = np.zeros((1, 4096), np.float16); xs.append(x)
    x = np.zeros((1, 4096), np.float16); x[0, :16] = 65504.0; x[0, 16:32] = -65504.0; x[0, 48:64] = 6e-8; x[0, 64] = -0.0
    xs.append(x)
    x = np.zeros((1, 4096), np.float16); x[0, ::16] = rng.standard_normal(256).astype(np.float16); xs.append(x)
    x = (rng.standard_normal((1, 4096)) * 1e-6).astype(np.float16); xs.append(x)          # subnormal range only
    x = np.full((1, 4096), 3.0, np.float16); xs.append(x)
    # keep |y| inside fp16 for the fp16-max case: small scales
    for k in ("scales_2nd", "scales_4b"):
        ps[0][k] = (ps[0][k].astype(np.float32) * 1e-3).astype(np.float16)
        pd[0][k].copy_(torch.from_numpy(ps[0][k]))
    xd = [torch.from_numpy(x).to(cuda) for x in xs]
    ys = [torch.full((256,), float("nan"), dtype=torch.float16, device=cuda) for _ in xs]
    ops.GemvChain([(a, pd[0], y, -1) for a, y in zip(xd, ys)]).run()
    torch.cuda.synchronize()
    assert torch.equal(ys[0], torch.zeros_like(ys[0]))
    for i, (x, y) in enumerate(zip(xs, ys)):
        assert torch.isfinite(y).all(), i
        _check(y.cpu().numpy(), x, ps[0], f"case {i}")


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_chain_random_dependency_graphs(cuda, seed):
    """Random DAGs over mixed shapes: a job reads either a fresh vector or the output of a random earlier job
    of the right width; siblings that read the same producer share its image; some producers are tiny (fewer
    tiles than CTAs, so most CTAs never publish them).  Every job is checked against the oracle applied to the
    input it actually read."""
    from mxq_b200 import ops
    rng = np.random.default_rng(100 + seed)
    widths = [256, 1024, 4096]
    n = 28
    shapes, deps = [], []
    for j in range(n):
        cands = [i for i in range(j) if True]
        if j > 0 and rng.random() < 0.7:
            d = int(rng.choice(cands))
            ic = shapes[d][0]
        else:
            d, ic = -1, int(rng.choice(widths))
        oc = int(rng.choice(widths + [32, 160]))
        if oc not in widths:                      # odd widths cannot feed another job
            pass
        shapes.append((oc, ic))
        deps.append(d)
    # producers must have a width another job can read: re-draw consumers of odd-width producers as fresh inputs
    for j in range(n):
        if deps[j] >= 0 and shapes[deps[j]][0] % 256:
            deps[j] = -1
            shapes[j] = (shapes[j][0], int(rng.choice(widths)))
    ps, pd = _mk(cuda, shapes, seed=300 + seed)
    for p, q in zip(ps, pd):                      # keep chains of several links inside fp16
        for k in ("scales_2nd", "scales_4b"):
            p[k] = (p[k].astype(np.float32) * 0.15).astype(np.float16)
            q[k].copy_(torch.from_numpy(p[k]))
    fresh = {w: torch.from_numpy(_outlier_x(1, w, seed=w + seed)[0]).to(cuda) for w in widths}
    ys = [torch.zeros(oc, dtype=torch.float16, device=cuda) for oc, _ in shapes]
    jobs = [((ys[deps[j]] if deps[j] >= 0 else fresh[shapes[j][1]]), pd[j], ys[j], deps[j]) for j in range(n)]
    chain = ops.GemvChain(jobs)
    for rep in range(2):
        for y in ys:
            y.fill_(float("nan"))
        chain.run()
        torch.cuda.synchronize()
        vals = [y.cpu().numpy() for y in ys]
        for j in range(n):
            xin = vals[deps[j]] if deps[j] >= 0 else fresh[shapes[j][1]].cpu().numpy()
            assert np.isfinite(vals[j]).all(), (rep, j)
            _check(vals[j], xin[None, :], ps[j], f"seed {seed} rep {rep} job {j} dep {deps[j]} {shapes[j]}")


def test_chain_longer_than_one_launch(cuda):
    from mxq_b200 import ops
    from mxq_b200 import _lib as L
    n = L.GEMV_CHAIN_MAX_JOBS + 5
    ps, pd = _mk(cuda, [(64, 256)] * 3, seed=3)
    x = _outlier_x(1, 256, seed=4)
    xd = torch.from_numpy(x).to(cuda)
    ys = [torch.zeros(1, 64, dtype=torch.float16, device=cuda) for _ in range(n)]
    chain = ops.GemvChain([(xd, pd[i % 3], ys[i], -1) for i in range(n)])
    assert len(chain._launches) == 2
    chain.run()
    torch.cuda.synchronize()
    for i in (0, 1, 2, n - 6, n - 1):
        _check(ys[i].cpu().numpy()[0], x, ps[i % 3], f"job {i}")


def test_chain_rejects_unsupported_shapes(cuda):
    from mxq_b200 import ops
    ps, pd = _mk(cuda, [(64, 320)])
    x = torch.zeros(320, dtype=torch.float16, device=cuda)
    y = torch.zeros(64, dtype=torch.float16, device=cuda)
    with pytest.raises(RuntimeError):
        ops.GemvChain([(x, pd[0], y, -1)])
    with pytest.raises(ValueError):
        ops.GemvChain([(x[:256], pd[0], y, -1)])
    ps, pd = _mk(cuda, [(40, 256)])                       # OC % 32 != 0
    with pytest.raises(RuntimeError):
        ops.GemvChain([(x[:256], pd[0], torch.zeros(40, dtype=torch.float16, device=cuda), -1)])


def test_chain_pdl_waits_for_the_previous_kernel(cuda):
    """MXQ_GEMV_CHAIN_PDL: a launch may start while the previous kernel of the stream is still running, but
    it must not read an activation vector before that kernel has completed.  Chain B reads what chain A
    writes (two launches, no in-chain dependency), chain A reads a vector that a copy kernel has just
    rewritten; eager back-to-back launches and a captured graph replayed with new inputs must both equal
    the unoverlapped result bit for bit."""
    from mxq_b200 import ops
    ic, mid, oc = 4096, 4096, 11008
    ps, pd = _mk(cuda, [(mid, ic), (mid, ic), (oc, mid), (oc, mid)], seed=11)
    x = torch.zeros(ic, dtype=torch.float16, device=cuda)
    xs = [torch.from_numpy(_outlier_x(1, ic, seed=40 + i)[0] * 0.05).to(cuda) for i in range(3)]
    h = [torch.zeros(mid, dtype=torch.float16, device=cuda) for _ in range(2)]
    y = [torch.zeros(oc, dtype=torch.float16, device=cuda) for _ in range(2)]
    A = ops.GemvChain([(x, pd[0], h[0], -1), (x, pd[1], h[1], -1)])
    B = ops.GemvChain([(h[0], pd[2], y[0], -1), (h[1], pd[3], y[1], -1)])

    def step(xi, pdl):
        x.copy_(xi)
        A.run(pdl=pdl)
        B.run(pdl=pdl)

    want = []
    for xi in xs:
        step(xi, False)
        torch.cuda.synchronize()
        want.append([t.clone() for t in h + y])
    for xi, w in zip(xs, want):                       # eager, overlapped
        step(xi, True)
        torch.cuda.synchronize()
        assert all(torch.equal(a, b) for a, b in zip(h + y, w))
    xin = torch.zeros_like(x)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        step(xin, True)
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for _ in range(3):                        # chain -> chain -> copy -> chain ... programmatic edges
                step(xin, True)
    for xi, w in zip(xs, want):
        xin.copy_(xi)
        for t in h + y:
            t.zero_()
        g.replay()
        torch.cuda.synchronize()
        assert all(torch.equal(a, b) for a, b in zip(h + y, w))
    _check(want[0][2].cpu().numpy(), want[0][0].cpu().numpy()[None], ps[2], "B[0] on A[0]'s output")
