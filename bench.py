#!/usr/bin/env python
"""bench.py -- Newton steps/s and ms per KKT-solve of the interior-point hot path at BASELINE.json's headline
configuration (config 3: synthetic nonconvex NLP, n = 4096 dense, 512 eq + 4096 ineq, fp64).

    python bench.py --gpus N --steps K --warmup W            # this repo's B200 path
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU arithmetic (oracle port)

One "step" = one inner iteration of the reference (pyipm.py:1714-1754): gradient/KKT residual, Lagrangian
Hessian, KKT formation, inertia-corrected factorisation (reghess), solve, nu rule, step rules, line search,
KKT conditions at the new point -- teacher-forced from a fixed state a few iterations into the real solve, so
every timed step does identical work.

N > 1 (torchrun, one rank per GPU): the Newton step at n = 4096 does not shard (SURVEY.md section 8e,
DESIGN.md "replicas only"): every rank steps an independent replica, no data-path collective, `scaling: weak`.

Prints ONE JSON line (rank 0).  See DESIGN.md section "Measurement" for every field.
"""
from __future__ import print_function

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = 'config3: synthetic nonconvex NLP n=4096, 512 eq + 4096 ineq, dense d2L, fp64 (problems.make_nlp())'
D3, M3, N3 = 4096, 512, 4096
K3 = D3 + 2 * N3 + M3
PRE_STEPS = 3   # real Newton steps taken from x0 before the teacher-forced state is snapshotted
# diagonal shift reghess settles on for the CPU sample state below (found once with the oracle itself: 10 eigvalsh
# calls, delta0 * 10^8); starting the sample at 2x this value makes every sampled step do the typical mid-solve
# work of two inertia tests (delta = 0 rejected, delta/2 accepted) instead of the 10-test discovery.
DELTA_SAMPLE = 1.4901161193847656
DEFAULT_FLAGS = 6     # tcgen05 int8 contractions (128x128 tiles) + speculative / abandoning reghess; 0 = fp64 DMMA contractions


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get('hbm_gbs', 6650.0), d.get('bf16_tflops', 1590.0), 'measured'
    return 6650.0, 1590.0, 'fallback'


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '20'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, universal_newlines=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                for name, col in (('hw_slowdown', 5), ('hw_thermal_slowdown', 6), ('sw_thermal_slowdown', 7),
                                  ('sw_power_cap', 8)):
                    if r[col].lower().startswith('active'):
                        reasons.add(name)
            except Exception:
                pass
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons),
                'samples': len(sm)}


# ------------------------------------------------------------------------------------------- reference arm
def oracle_sample_step(D=1024, M=128, N=1024, seed=7):
    """One Newton step of the CPU oracle on a bounded sample of the workload: the SAME problem family at
    n = 1024 (K = 3200 instead of 12800), teacher-forced a few iterations in with delta already active so the
    step does the typical two inertia tests (delta = 0 fails, delta/2 passes) + one LU.  A config-3-size step is
    >= 4 minutes (eigvalsh(12800) = 235 s + LU 17 s on 8 cores) and several times that when the
    delta loop escalates, so it cannot be a bench step; the measured n=1024 time is scaled by the cubic
    work law (K3/K)^3 = 64 of the dense eigen/LU kernels that make up > 95 % of it."""
    from oracle.pyipm_numpy import OracleIPM
    from pyipm_b200 import problems
    prob = problems.make_nlp(D=D, M=M, N=N, seed=seed)
    o = OracleIPM(x0=prob.x0.copy(), verbosity=-1, **prob.callables())
    o.nvar = D
    o.compile()
    rng = np.random.default_rng(seed)
    x = prob.x0.copy()
    s = np.maximum(prob.ci(x), 1e-4)
    lda = np.concatenate([0.1 * rng.standard_normal(M), 0.2 / s])
    o.mu_host = 0.2
    o.mu_dev = np.float64(0.2)
    o.nu_host = 10.0
    o.nu_dev = np.float64(10.0)
    o.signal = 0
    return o, prob, (x, s, lda)


STATE_FILE = os.path.join(ROOT, 'tests', 'golden', 'c3_state_after3.npz')


def load_c3_state():
    """The teacher-forced config-3 state both arms step from: the iterate after PRE_STEPS real Newton steps from x0
    (tools/make_c3_state.py, produced once on a B200 and committed: the reference arm has no GPU to produce it with)."""
    g = np.load(STATE_FILE)
    return (g['x'], g['s'], g['lda'], float(g['mu']), float(g['nu']), float(g['delta']), float(g['mu_host']))


def oracle_c3_step():
    """ONE real config-3 Newton step of the CPU oracle (pyipm.py:1714-1754 restated) from the committed state: the same
    problem, the same iterate and the same (mu, nu, delta) as the B200 arm's timed step.  delta > 0 on entry, so reghess
    does what it does in the middle of the solve: eigvalsh(12800) at delta = 0 (fails), eigvalsh at delta / 2 (passes),
    then one LU solve of the 12800^2 system."""
    from oracle.pyipm_numpy import OracleIPM
    from pyipm_b200 import problems
    prob = problems.make_nlp(D3, M3, N3)
    x, s, lda, mu, nu, delta, mu_host = load_c3_state()
    o = OracleIPM(x0=prob.x0.copy(), verbosity=-1, mu=mu, **prob.callables())
    o.nvar = D3
    o.compile()
    o.mu_host, o.mu_dev, o.nu_host, o.nu_dev, o.signal = mu_host, np.float64(mu), nu, np.float64(nu), 0
    o.delta = np.float64(delta)
    o.timers = {}
    tr = []
    o.trace = tr
    t0 = time.perf_counter()
    with np.errstate(all='ignore'):
        o.newton_step(x.copy(), s.copy(), lda.copy())
    dt = time.perf_counter() - t0
    return o, tr[0], dt


def run_reference(args):
    """The reference's own CPU arithmetic on the SAME configuration: one real config-3 step (about 5-10 minutes on 16
    cores: two eigvalsh(12800) + one LU).  --steps / --warmup beyond one step are ignored -- the line says steps = 1,
    warmup = 0 -- because K steps of this size do not fit any bench window; the n = 1024 sample of the same family is kept
    as a labelled extra (`sample_n1024`)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    os.environ.setdefault('OPENBLAS_NUM_THREADS', str(os.cpu_count()))
    cores = os.cpu_count()
    extra = None
    if not args.ref_skip_sample:
        o1, p1, (x1, s1, l1) = oracle_sample_step()
        o1.delta = np.float64(2.0 * DELTA_SAMPLE)
        o1.timers = {}
        t0 = time.perf_counter()
        with np.errstate(all='ignore'):
            o1.newton_step(x1.copy(), s1.copy(), l1.copy())
        t1 = time.perf_counter() - t0
        K1 = p1.nvar + 2 * p1.nineq + p1.neq
        extra = {'n': p1.nvar, 'K': K1, 'step_s': t1, 'n_eigvalsh': o1.last_reg['n_eig'],
                 'cubic_extrapolation_to_config3_s': t1 * (float(K3) / K1) ** 3,
                 'note': 'same NLP family at n=1024; NOT the headline number, kept to show the cubic law'}
    if args.ref_test_size:
        # contract test only (tests/test_bench_contract_cpu.py): the same code path on a toy instance of the family
        o, p1, (x1, s1, l1) = oracle_sample_step(D=args.ref_test_size, M=args.ref_test_size // 8, N=args.ref_test_size)
        o.delta = np.float64(2.0 * DELTA_SAMPLE)
        o.timers = {}
        tr = []
        o.trace = tr
        t0 = time.perf_counter()
        with np.errstate(all='ignore'):
            o.newton_step(x1.copy(), s1.copy(), l1.copy())
        dt = time.perf_counter() - t0
        st = tr[0]
    else:
        o, st, dt = oracle_c3_step()
    value = 1.0 / dt
    sample = ('ONE full config-3 oracle Newton step (K=%d) from tests/golden/c3_state_after3.npz: %.1f s, %d eigvalsh(12800) '
              '+ 1 LU; requested steps=%d warmup=%d ignored beyond one step' % (K3, dt, st['reg']['n_eig'], args.steps, args.warmup))
    line = {
        'impl': 'reference', 'metric': 'newton_steps_per_sec', 'value': value, 'unit': 'steps/s', 'n_gpus': args.gpus,
        'steps': 1, 'warmup': 0, 'steps_requested': args.steps, 'warmup_requested': args.warmup,
        'ms_per_step': 1e3 * dt, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': WORKLOAD + (' -- CONTRACT TEST AT REDUCED SIZE D=%d, not a measurement' % args.ref_test_size
                                           if args.ref_test_size else ''), 'D': D3, 'M': M3, 'N': N3, 'K_full': K3,
                   'state': 'teacher-forced from the state after %d real Newton steps from x0 (committed fixture)' % PRE_STEPS},
        'cpu_baseline': {'value': value, 'unit': 'steps/s', 'cores': cores, 'kind': 'port', 'sample': sample,
                         'same_config': True, 'split_s': {k: v for k, v in o.timers.items() if k != 'steps'}},
        'e2e': {'value': value, 'unit': 'steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'ms_per_kkt_solve': 1e3 * (o.timers.get('reghess', 0.0) + o.timers.get('solve', 0.0)),
        'reghess': {'n_eigvalsh': st['reg']['n_eig'], 'delta': st['delta'], 'n_backtracks': st['search']['n_backtracks']},
        'sample_n1024': extra,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------- config 4 (multi-GPU LDL^T)
def measure_c4(world, rank, local, n, steps, warmup, fp64_peak_tf=None):
    """BASELINE config 4: dense symmetric quasi-definite KKT of order n (default 16384), 2-D block-cyclic LDL^T over the GPUs
    of one node (pyipm_b200/dist_ldlt.py: look-ahead pipeline, NCCL panel broadcasts), 8 right-hand sides.  STRONG scaling:
    the matrix is the same for every world size.  Needs an initialised process group when world > 1."""
    import torch
    import torch.distributed as dist
    from pyipm_b200.dist_ldlt import BlockCyclicLDLT, CudaTileOps, choose_grid
    m = n // 8
    nh = n - m
    g = torch.Generator(device='cuda')
    g.manual_seed(16384)
    W = torch.randn(nh, nh, dtype=torch.float64, device='cuda', generator=g)
    K = torch.zeros(n, n, dtype=torch.float64, device='cuda')
    K[:nh, :nh] = W @ W.t() / nh
    del W
    K[:nh, :nh].diagonal().add_(10.0 ** (8.0 * torch.rand(nh, dtype=torch.float64, device='cuda', generator=g) - 4.0))
    J = torch.randn(nh, m, dtype=torch.float64, device='cuda', generator=g)
    K[:nh, nh:] = J
    K[nh:, :nh] = J.t()
    K[nh:, nh:].diagonal().fill_(-1e-8)
    rhs = torch.randn(8, n, dtype=torch.float64, device='cuda', generator=g)
    grid = choose_grid(world)
    native = None
    if world == 1:
        # strong-scaling base: the best single-GPU implementation, i.e. the engine's own CUDA-graph factorisation (tcgen05
        # trailing updates, block-256 solves) through the generic C-ABI slot, matrix and right-hand sides resident
        import ctypes as C
        from pyipm_b200 import _lib
        lib = _lib.load()
        Fn = _lib.DenseLDLT(n, device=local, stream=_lib.torch_stream_handle(torch.device('cuda', local)))
        tf_n, ts_n = [], []
        for it in range(warmup + steps):
            Xn = rhs.clone()
            torch.cuda.synchronize()
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            ine = (C.c_int * 3)()
            rc = C.c_double()
            e0.record()
            _lib.check(lib.b200ipm_ldlt_factor(Fn.h, C.c_void_p(K.data_ptr()), n, 1, ine, C.byref(rc)))
            e1.record()
            _lib.check(lib.b200ipm_ldlt_solve(Fn.h, C.c_void_p(Xn.data_ptr()), 8, 1, 1))
            e2.record()
            torch.cuda.synchronize()
            if it >= warmup:
                tf_n.append(e0.elapsed_time(e1))
                ts_n.append(e1.elapsed_time(e2))
        Rn = rhs - Xn @ K
        native = {'factor_ms': float(np.mean(tf_n)), 'solve_ms_8rhs_1refine': float(np.mean(ts_n)), 'inertia': list(ine),
                  'scaled_residual_inf': float(Rn.abs().max() / (K.abs().max() * Xn.abs().max())),
                  'factor_tflops': n ** 3 / 3.0 / float(np.mean(tf_n)) * 1e-9,
                  'note': 'b200ipm_ldlt_factor / b200ipm_ldlt_solve (one CUDA graph, tcgen05 trailing updates), incl. the '
                          'device copy + symmetrisation of the input'}
        Fn.close()
        del Fn, Xn, Rn
    F = BlockCyclicLDLT(n, grid, CudaTileOps(local), block=256)
    F.load_device(K)
    times_f, times_s = [], []
    inertia = None
    for it in range(warmup + steps):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        inertia = F.factor()
        e1.record()
        X = F.solve_device(rhs, nrefine=1)
        e2.record()
        torch.cuda.synchronize()
        tf, ts = e0.elapsed_time(e1), e1.elapsed_time(e2)
        if world > 1:
            t = torch.tensor([tf, ts], device='cuda', dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            tf, ts = float(t[0]), float(t[1])
        if it >= warmup:
            times_f.append(tf)
            times_s.append(ts)
    R = rhs - F.matvec(X)
    resid = float(R.abs().max() / (K.abs().max() * X.abs().max()))
    tf, ts = float(np.mean(times_f)), float(np.mean(times_s))
    tfl = n ** 3 / 3.0 / tf * 1e-9
    rec = {'workload': 'config4: dense symmetric quasi-definite KKT order %d (%d + %d), 8 RHS, block-column-cyclic LDL^T, grid '
                       '%dx%d, block 256, look-ahead 1, one broadcast per block column' % (n, nh, m, grid[0], grid[1]),
           'n_gpus': world, 'scaling': 'strong', 'factor_ms': tf, 'solve_ms_8rhs_1refine': ts,
           'factor_tflops_aggregate': tfl, 'inertia': list(inertia), 'inertia_expected': [nh, m, 0],
           'scaled_residual_inf': resid, 'steps': steps, 'warmup': warmup}
    if native is not None:
        # N = 1: the headline numbers of the record are the single-GPU implementation; the distributed pipeline run on
        # one rank (no collectives) is kept beside it
        rec['pipeline_on_one_rank'] = {'factor_ms': tf, 'solve_ms_8rhs_1refine': ts, 'factor_tflops': tfl}
        rec['implementation'] = 'single-GPU CUDA-graph factorisation (ldlt.cuh) -- the strong-scaling base'
        rec['factor_ms'], rec['solve_ms_8rhs_1refine'] = native['factor_ms'], native['solve_ms_8rhs_1refine']
        rec['factor_tflops_aggregate'] = tfl = native['factor_tflops']
        rec['native'] = native
    else:
        rec['implementation'] = 'pyipm_b200/dist_ldlt.py block-column-cyclic pipeline over NCCL'
    # HBM leg of the record: every triangular solve streams the factor twice (forward + backward: n^2 * 8 bytes), the
    # refinement mat-vec streams each rank's share of the original matrix; replicated solves => per-GPU traffic
    nsolve = 8 * 2
    sol_bytes = nsolve * float(n) * n * 8.0 + 8 * float(n) * n * 8.0 / world
    rec['solve_bytes_per_gpu'] = sol_bytes
    rec['solve_gbs_per_gpu'] = sol_bytes / (rec['solve_ms_8rhs_1refine'] * 1e-3) * 1e-9
    try:
        rec['solve_frac_of_hbm_peak'] = rec['solve_gbs_per_gpu'] / measured_peaks()[0]
    except Exception:
        pass
    if fp64_peak_tf:
        rec['fp64_peak_tf_per_gpu'] = fp64_peak_tf
        rec['frac_of_aggregate_fp64_peak'] = tfl / (world * fp64_peak_tf)
    del F, K
    torch.cuda.empty_cache()
    return rec


def run_c4(args):
    """`--workload c4`: config 4 alone (its own JSON line; not the headline metric)."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    rec = measure_c4(world, rank, local, args.c4_n, args.steps, args.warmup)
    if rank == 0:
        line = {'metric': 'kkt_factor_solve_ms', 'value': rec['factor_ms'] + rec['solve_ms_8rhs_1refine'], 'unit': 'ms',
                'n_gpus': world, 'higher_is_better': False, 'steps': args.steps, 'warmup': args.warmup, 'scaling': 'strong',
                'dtype': 'f64', 'data': 'synthetic', 'config': {'workload': rec['workload']}}
        line.update(rec)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


# ------------------------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    from pyipm_b200 import _lib, problems

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device (the B200 arm has no CPU fallback; use --impl reference)')
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    hbm_gbs, bf16_tf, peak_kind = measured_peaks()
    prob = problems.make_nlp(D3, M3, N3)       # same seed on every rank: independent replicas of config 3
    stream = _lib.torch_stream_handle()
    eng = _lib.Engine(D3, M3, N3, _lib.default_params(flags=args.flags), device=local, stream=stream)
    eng.bind(prob)
    eng.set_state(prob.x0, np.ones(N3), np.zeros(M3 + N3), 0.2, 10.0, 0.0)
    eng.set_mu_host(0.2)
    eng.init_slack()
    eng.init_lambda()
    pre = [eng.newton_step().asdict() for _ in range(PRE_STEPS)]
    eng.state_save()
    x_h, s_h, l_h, mu, nu, delta = eng.get_state()
    pin = [torch.from_numpy(a).pin_memory() for a in (x_h, s_h, l_h)]
    out = [torch.empty_like(t).pin_memory() for t in pin]

    def step_resident():
        eng.state_restore()
        return eng.newton_step()

    def step_e2e():
        # host buffers in, host buffers out: the call a user of the C ABI makes per iteration
        _lib.check(eng.lib.b200ipm_set_state(eng.h, pin[0].data_ptr(), pin[1].data_ptr(), pin[2].data_ptr(), mu, nu, delta))
        info = eng.newton_step()
        _lib.check(eng.lib.b200ipm_get_state(eng.h, out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), None, None, None))
        return info

    # fp64 contraction peak on THIS box (MEASURED_PEAKS.json only carries bf16 + HBM): cuBLAS DGEMM 4096^3
    a = torch.randn(4096, 4096, dtype=torch.float64, device='cuda')
    b = torch.randn(4096, 4096, dtype=torch.float64, device='cuda')
    torch.matmul(a, b)
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    fp64_peak_tf = 2 * 4096 ** 3 / best * 1e-9
    del a, b
    # int8 tensor-core throughput of the vendor library on THIS box (cuBLASLt via torch._int_mm), for the tcgen05 leg
    int8_peak_tops = None
    try:
        ai = torch.randint(-64, 64, (8192, 8192), dtype=torch.int8, device='cuda')
        bi = torch.randint(-64, 64, (8192, 8192), dtype=torch.int8, device='cuda')
        torch._int_mm(ai, bi)
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); torch._int_mm(ai, bi); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        int8_peak_tops = 2 * 8192 ** 3 / best * 1e-9
        del ai, bi
    except Exception:
        pass

    def timed(fn, steps, warmup):
        # nvidia-smi needs ~0.1 s before its first sample and K steps last well under that: the sampler starts before the
        # warm-up, and extra untimed warm-up steps keep the GPU under the same load until samples are flowing, so that
        # the clock record covers the load of the timed region (which follows immediately)
        sampler = ClockSampler(local)
        sampler.start()
        for _ in range(warmup):
            fn()
        t_wait = time.time()
        while len(sampler.rows) < 2 and time.time() - t_wait < 2.0:
            fn()
        barrier()
        n0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        infos = [fn() for _ in range(steps)]
        e1.record()
        barrier()
        clocks = sampler.stop()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device='cuda', dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, infos, _lib.launch_count() - n0, clocks

    ms_res, infos, launches, clocks = timed(step_resident, args.steps, args.warmup)
    ms_e2e, infos_e, _, _ = timed(step_e2e, args.steps, max(args.warmup, 3))
    # real-trajectory leg: consecutive TRUE Newton steps from x0 (no state restore: cold certificate vectors, second-order
    # corrections and delta discovery included), timed per step on the device
    traj = None
    if rank == 0 and args.traj_steps > 0:
        cold = None
        for trip in range(2):     # first trip: cold (lazily created workspaces, graph captures); second trip: the record
            eng.set_state(prob.x0, np.ones(N3), np.zeros(M3 + N3), 0.2, 10.0, 0.0)
            eng.set_mu_host(0.2)
            eng.init_slack()
            eng.init_lambda()
            eng.sync()
            tinfo = []
            t0 = time.perf_counter()
            for _ in range(args.traj_steps):
                tinfo.append(eng.newton_step())
            wall = (time.perf_counter() - t0) * 1e3
            ms = [float(i.ms_total) for i in tinfo]
            if trip == 0:
                cold = {'ms_mean': float(np.mean(ms)), 'ms_max': float(np.max(ms)), 'ms_per_step_list': [round(v, 3) for v in ms],
                        'note': 'first trip over the same trajectory in this process: includes the one-time allocation + CUDA-graph '
                                'capture of workspaces created at first use (second-order correction, strict re-factorisation)'}
        traj = {'steps': len(tinfo), 'cold_first_trip': cold, 'ms_mean': float(np.mean(ms)), 'ms_median': float(np.median(ms)), 'ms_max': float(np.max(ms)),
                'ms_min': float(np.min(ms)), 'wall_ms_per_step': wall / len(tinfo),
                'ms_per_step_list': [round(v, 3) for v in ms],
                'n_soc_tried': int(sum(i.soc_tried for i in tinfo)), 'n_soc_accepted': int(sum(i.soc_accepted for i in tinfo)),
                'n_cert_used': int(sum(i.cert_used for i in tinfo)), 'n_speculated': int(sum(i.n_spec for i in tinfo)),
                'factorisations_reference_equivalent': [int(i.n_factor) for i in tinfo],
                'factorisations_physical': [int(i.n_factor_phys) for i in tinfo],
                'n_backtracks': [int(i.n_backtracks) for i in tinfo],
                'soc_tried': [int(i.soc_tried) for i in tinfo], 'cert_used': [int(i.cert_used) for i in tinfo],
                'phase_ms': {k: [round(float(getattr(i, k)), 3) for i in tinfo] for k in
                             ('ms_eval', 'ms_assemble', 'ms_factor', 'ms_solve', 'ms_search')},
                'kkt_norm_last': [float(v) for v in tinfo[-1].kkt_norm],
                'note': 'niter=1-style inner iterations from x0 at fixed mu = 0.2 (no barrier update), step 1 discovers delta '
                        '(10 inertia tests), later steps speculate; ms from CUDA events inside b200ipm_newton_step'}
    state_check = None
    if os.path.exists(STATE_FILE):
        xf, sf, lf, muf, nuf, df_, _ = load_c3_state()
        state_check = {'fixture': 'tests/golden/c3_state_after3.npz',
                       'max_rel_diff_x': float(np.max(np.abs(xf - x_h)) / np.max(np.abs(x_h))),
                       'max_rel_diff_lda': float(np.max(np.abs(lf - l_h)) / np.max(np.abs(l_h))),
                       'delta_equal': bool(df_ == delta), 'nu_rel_diff': float(abs(nuf - nu) / abs(nu))}
    per_step = ms_res / args.steps
    value = world * args.steps / (ms_res * 1e-3)
    e2e_value = world * args.steps / (ms_e2e * 1e-3)
    last = infos[-1].asdict()
    kkt_ms = float(np.mean([i.ms_factor + i.ms_solve for i in infos]))
    nfac = float(np.mean([i.n_factor for i in infos]))
    nfac_phys = float(np.mean([i.n_factor_phys for i in infos]))

    line = {
        'metric': 'newton_steps_per_sec', 'value': value, 'unit': 'steps/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': per_step, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'D': D3, 'M': M3, 'N': N3, 'K_full': K3, 'K_condensed': D3 + M3,
                   'parallelism': 'replicas x%d (one independent Newton step per GPU, no collective)' % world,
                   'state': 'teacher-forced from the state after %d real Newton steps from x0' % PRE_STEPS,
                   'factorisations_per_step_reference_equivalent': nfac, 'factorisations_per_step_physical': nfac_phys,
                   'cert_used_per_step': float(np.mean([i.cert_used for i in infos])),
                   'refinement_sweeps': 2, 'engine_flags': args.flags,
                   'contractions': 'tcgen05 int8 error-free split' if args.flags & 2 else 'fp64 DMMA',
                   'reghess': ('sequential' if args.flags & 1 else 'delta=0 test in the background, candidate in the foreground'),
                   'l2': 'working set (J 151 MB, Vt/Gt/Q 134 MB each, KKT 170 MB) exceeds the 126 MB L2; no flush'},
        'clocks': clocks,
        'e2e': {'value': e2e_value, 'unit': 'steps/s', 'h2d_bytes_per_step': int(8 * (D3 + N3 + M3 + N3)),
                'd2h_bytes_per_step': int(8 * (D3 + N3 + M3 + N3)), 'ms_per_step': ms_e2e / args.steps,
                'note': 'b200ipm_set_state(host) + b200ipm_newton_step + b200ipm_get_state(host) per step; problem '
                        'matrices are bound once (they are the problem definition, like the reference\'s compiled '
                        'functions), the per-step inputs/outputs are the iterate (x, s, lda)'},
        'gpu_launches': int(launches),
        'ms_per_kkt_solve': kkt_ms,
        'kkt_residual_inf': last['resid'],
        'phase_ms': {k: float(np.mean([getattr(i, k) for i in infos])) for k in
                     ('ms_eval', 'ms_assemble', 'ms_factor', 'ms_solve', 'ms_search', 'ms_total')},
        'trajectory': traj, 'state_check': state_check,
    }
    # config 4 (the path that DOES shard): every rank takes part; the record rides on the same JSON line so that the driver's
    # 1/2/4/8-GPU runs capture its strong scaling
    c4 = None
    if args.c4_n > 0:
        try:
            c4 = measure_c4(world, rank, local, args.c4_n, 3, 3, fp64_peak_tf)   # 3 warm-ups: the caching allocator settles
        except Exception as exc:
            c4 = {'error': repr(exc)}
    line['c4'] = c4
    if rank == 0:
        # roofline legs: the dominant kernel (fp64 DMMA contraction) and the HBM-bound residual GEMV, each timed
        # alone with CUDA events on the launching stream
        eng.state_restore()
        kern = {}
        for which, name, kind in ((1, 'hess_syrk', 'flop'), (2, 'condense_syrk', 'flop'), (3, 'ldlt_factor', 'flop'),
                                  (0, 'residual_gemv', 'byte'), (5, 'jt_gemv', 'byte'), (4, 'ldlt_solve', 'byte'),
                                  (6, 'hess_syrk_tcgen05', 'flop'), (7, 'condense_syrk_tcgen05', 'flop')):
            ms, work = eng.profile_kernel(which, reps=5)
            kern[name] = {'ms': ms, ('tflops' if kind == 'flop' else 'gbs'): work / ms * (1e-9 if kind == 'flop' else 1e-6),
                          'work': work}
        traffic = {}
        tpath = os.path.join(ROOT, 'profiles', 'r2_traffic.json')
        if os.path.exists(tpath):
            traffic = json.load(open(tpath))
        tc_on = bool(args.flags & 2)
        ph = line['phase_ms']
        # dominant kernel family of the step: the LDL^T factorisation of the condensed KKT matrix (one CUDA graph).
        # Algorithmic work Kc^3 / 3 over the fp64 DGEMM rate measured in this run.
        fac = kern['ldlt_factor']
        line['roofline'] = {
            'kernel': 'ldlt_factor: tile-pivoted LDL^T of the condensed KKT matrix (order %d), one CUDA graph (tile / mini / '
                      'panel / DMMA + tcgen05 trailing updates)' % (D3 + M3),
            'bound': 'tensor', 'achieved': fac['tflops'], 'peak': fp64_peak_tf, 'unit': 'TFLOP/s',
            'frac': fac['tflops'] / fp64_peak_tf, 'ms': fac['ms'], 'flops_per_launch': fac['work'],
            'share_of_step': ph['ms_factor'] / ph['ms_total'] if ph['ms_total'] else None,
            'traffic': traffic.get('ldlt_factor'), 'traffic_source': traffic.get('source'),
            'peak_source': 'cuBLAS DGEMM 4096^3 measured in this run = %.1f TF/s (fp64 tensor pipe; MEASURED_PEAKS.json has '
                           'no fp64 entry: its bf16 figure %.0f TF/s (%s) is a different pipe)' % (fp64_peak_tf, bf16_tf, peak_kind)}
        if tc_on:
            ms8, ops8 = eng.profile_kernel(8, reps=5)
            ach = ops8 / ms8 * 1e-9
            peak = 2.0 * bf16_tf     # dense int8 runs at twice the bf16 rate on this part; bf16_tf is the MEASURED burst figure
            kern['hess_syrk_tcgen05_kernel_only'] = {'ms': ms8, 'int8_tops': ach, 'work': ops8}
            full = kern['hess_syrk_tcgen05']
            alg = full['work']     # fp64 FLOPs of the product itself (D^2 (M+N))
            line['roofline_tensor'] = {
                'kernel': 'oz_syrk_kernel<128,7> (Lagrangian-Hessian contraction Ut diag(lda_e) Ut\' + Vt diag(lda_i) Vt\' as 28 '
                          'exact int8 slice-pair products on tcgen05.mma.kind::i8, int32 TMEM accumulators, fp64 recombination)',
                'bound': 'tensor', 'unit': 'TFLOP/s',
                'issue_rate_int8_tops': ach, 'issue_rate_peak': peak, 'issue_rate_frac': ach / peak,
                'frac_of_cublaslt_int8': (ach / int8_peak_tops) if int8_peak_tops else None,
                'algorithmic_fp64_flops_per_launch': alg, 'int8_ops_per_launch': ops8,
                'algorithmic_tflops_kernel_only': alg / ms8 * 1e-9,
                'algorithmic_frac_of_int8_pipe': (alg / ms8 * 1e-9) / peak,
                'fp64_equivalent_tflops_incl_slicing': full['tflops'], 'fp64_dgemm_peak': fp64_peak_tf,
                'fp64_equivalent_vs_dgemm': full['tflops'] / fp64_peak_tf,
                'dmma_kernel_tflops': kern['hess_syrk']['tflops'],
                'share_of_step': float(np.mean([i.ms_hess_kernel for i in infos])) / ph['ms_total'] if ph['ms_total'] else None,
                'traffic': traffic.get('oz_syrk_kernel'), 'traffic_source': traffic.get('source'),
                'peak_source': ('2 x MEASURED_PEAKS.json bf16_tflops (%s; dense int8 = twice the bf16 rate) = %.0f TOP/s; cuBLASLt '
                                'int8 GEMM 8192^3 (torch._int_mm) measured in this run = %.0f TOP/s'
                                % (peak_kind, 2.0 * bf16_tf, int8_peak_tops or 0.0)),
                'note': 'issue_rate_* count the int8 multiply-adds the error-free split issues (28 slice pairs); algorithmic_* '
                        'count the fp64 FLOPs of the product (SURVEY 8d) -- tcgen05 has no f64 kind, so 28 int8 products '
                        'buy one fp64 product'}
        res = kern['residual_gemv']
        sol = kern['ldlt_solve']
        line['roofline_hbm'] = {'kernel': 'gemv_n_kernel (g_x = df - J*lda, fused KKT norm)', 'bound': 'hbm',
                                'achieved': res['gbs'], 'peak': hbm_gbs, 'unit': 'GB/s', 'frac': res['gbs'] / hbm_gbs,
                                'traffic': traffic.get('gemv_n_kernel'), 'traffic_source': traffic.get('source'),
                                'peak_source': 'MEASURED_PEAKS.json hbm_gbs (%s)' % peak_kind,
                                'bytes_per_launch': res['work'],
                                'triangular_solve': {'kernel': 'ldlt_fwd256_kernel + ldlt_bwd256_kernel: block-256 cluster solves, one right-hand side '
                                                               '(B200IPM_SOLVE256=0: the 64-row chain ldlt_fwd_kernel + ldlt_bwd_kernel)',
                                                     'ms': sol['ms'], 'achieved': sol['gbs'], 'frac': sol['gbs'] / hbm_gbs,
                                                     'bytes_per_launch': sol['work']}}
        line['kernels'] = kern
        # "KKT-residual match vs CPU ref": one teacher-forced Newton step of a small instance of the same NLP
        # family on this GPU against the CPU oracle (the full parity suite is tests/test_gpu_engine.py)
        try:
            from oracle.pyipm_numpy import OracleIPM
            sp = problems.make_nlp(D=192, M=32, N=192, seed=3)
            tr = []
            oq = OracleIPM(x0=sp.x0.copy(), verbosity=-1, niter=1, miter=2, trace=tr, **sp.callables())
            with np.errstate(all='ignore'):
                oq.solve()
            e2 = _lib.Engine(sp.nvar, sp.neq, sp.nineq, _lib.default_params(), device=local)
            e2.bind(sp)
            stq = tr[1]
            e2.set_state(stq['x'], stq['s'], stq['lda'], stq['mu'], tr[0]['nu_after'], tr[0]['delta'])
            e2.set_mu_host(stq['mu_host'])
            dzq, iq = e2.direction()
            e2.set_state(stq['x'], stq['s'], stq['lda'], stq['mu'], tr[0]['nu_after'], tr[0]['delta'])
            i2 = e2.newton_step()
            line['parity_check'] = {
                'problem': 'nlp D=192 M=32 N=192, step 2 of the oracle trajectory',
                'dz_rel_err_inf': float(np.max(np.abs(dzq - stq['dz'])) / np.max(np.abs(stq['dz']))),
                'same_delta': bool(iq.delta == stq['delta']), 'same_n_factor': bool(iq.n_factor == stq['reg']['n_eig']),
                'kkt_norms_gpu': list(i2.kkt_norm), 'kkt_norms_oracle': [float(v) for v in stq['kkt_norms']]}
            e2.close()
        except Exception as exc:   # the parity field is informative; never let it break the bench line
            line['parity_check'] = {'error': repr(exc)}
        if world == 1 and not args.no_cpu_baseline:
            os.environ.setdefault('OPENBLAS_NUM_THREADS', str(os.cpu_count()))
            o, sprob, (x, s, lda) = oracle_sample_step()
            K = sprob.nvar + 2 * sprob.nineq + sprob.neq
            scale = (float(K3) / K) ** 3
            with np.errstate(all='ignore'):
                ts = []
                for _ in range(3):
                    o.delta = np.float64(2.0 * DELTA_SAMPLE)
                    o.timers = {}
                    t0 = time.perf_counter()
                    o.newton_step(x.copy(), s.copy(), lda.copy())
                    ts.append(time.perf_counter() - t0)
                t_sample = float(np.median(ts))
            line['cpu_baseline'] = {
                'value': 1.0 / (t_sample * scale), 'unit': 'steps/s', 'cores': os.cpu_count(), 'kind': 'port',
                'same_config': False, 'scaled_from_n': sprob.nvar, 'scale_factor': scale,
                'sample': 'BOUNDED SAMPLE, extrapolated: oracle Newton step on the same NLP family at n=1024 (K=%d): median of 3 '
                          '= %.2f s (%d eigvalsh + 1 LU each), scaled to config 3 by (K3/K)^3 = %.0f.  The same-configuration '
                          'measurement is `bench.py --impl reference` (one real config-3 step, ~2 min on 16 cores; '
                          'profiles/r2_bench_reference.json)' % (K, t_sample, o.last_reg['n_eig'], scale),
                'split_s': {k: v for k, v in o.timers.items() if k != 'steps'}}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--ref-skip-sample', action='store_true', help='reference arm: skip the n=1024 extra')
    ap.add_argument('--ref-test-size', type=int, default=0, help=argparse.SUPPRESS)   # contract test: toy instance
    ap.add_argument('--traj-steps', type=int, default=12, help='real-trajectory leg: consecutive Newton steps from x0')
    ap.add_argument('--flags', type=int, default=DEFAULT_FLAGS,
                    help='b200ipm_params.flags: 1 no speculative reghess, 2 tcgen05 int8 SYRKs, (v << 2) tcgen05 tile variant')
    ap.add_argument('--workload', default='c3', choices=['c3', 'c4'])
    ap.add_argument('--c4-n', type=int, default=16384, help='order of the config-4 matrix (0: skip the c4 sub-record)')
    args = ap.parse_args()
    if args.workload == 'c4':
        return run_c4(args)
    if args.impl == 'reference':
        return run_reference(args)    # ONE real config-3 step (minutes); see run_reference
    args.warmup = max(args.warmup, 3)
    return run_b200(args)


if __name__ == '__main__':
    sys.exit(main())
