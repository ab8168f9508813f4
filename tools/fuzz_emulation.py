"""Randomized differential campaign of the not-yet-measured kernels on the SIMT / tcgen05 emulation (tests/simt_emu):
random shapes incl. row / channel tails against tests/fake_backend.py.   python tools/fuzz_emulation.py SEED TRIALS"""
import sys, os, subprocess, random, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tests.test_simt_emulation as T
from tests.fake_backend import FakeBackend
from torch_semantic_segmentation_b200 import _lib
so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'build', 'host_emu_fuzz.so')
cmd = ['g++', '-std=c++20', '-O1', '-ffp-contract=off', '-DTSS_HOST_EMU', '-x', 'c++', '-shared', '-fPIC', '-pthread',
       '-I', T.EMU, '-I', T.CSRC] + [os.path.join(T.CSRC, f) for f in T.SOURCES] + [os.path.join(T.EMU, 'emu_runtime.cpp'), '-o', so]
subprocess.check_call(cmd)
emu = T.EmulatedBackend(so)
rel = T.rel
random.seed(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
g = torch.Generator().manual_seed(random.randrange(1 << 30))
nh = lambda N, C, H, W, dt: torch.randn(N, H, W, C, generator=g).to(dt).permute(0, 3, 1, 2)
par = lambda C: (torch.randn(C, generator=g) * 0.2, torch.rand(C, generator=g) + 0.5, torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.3)
worst = {}
def note(k, v, lim):
    worst[k] = max(worst.get(k, 0), v)
    assert v < lim, (k, v)
for trial in range(int(sys.argv[2]) if len(sys.argv) > 2 else 12):
    # ---- pw fused backward
    N, H, W = random.randint(1, 3), random.randint(1, 11), random.randint(1, 13)
    K = 16 * random.randint(1, 8); Nc = 8 * random.randint(1, 96); relu = random.randint(0, 1); link = random.randint(0, 1)
    M = N * H * W; dt = torch.bfloat16
    dz, y, yp = nh(N, Nc, H, W, dt), nh(N, Nc, H, W, dt), nh(N, K, H, W, dt)
    mean, rstd, gamma, beta = par(Nc); pm, pr, pg, pb = par(K)
    wpT = (torch.randn(K, Nc, generator=g) / Nc ** 0.5).to(dt)
    fake = FakeBackend(); sums = torch.zeros(2 * Nc)
    fake.call('tss_bn_bwd_reduce', dict(dz=dz, z=None, y=y, mean=mean, rstd=rstd, gamma=gamma, beta=beta, sums=sums, M=M, C=Nc, lddz=Nc, ldz=0, ldy=Nc, flags=relu, dtype=1))
    outs = {}
    for name, be in (('ref', fake), ('emu', emu)):
        dy = torch.zeros(N, H, W, Nc, dtype=dt).permute(0, 3, 1, 2); dx = torch.zeros(N, H, W, K, dtype=dt).permute(0, 3, 1, 2)
        dga, dbe, ps = torch.ones(Nc), torch.ones(Nc), torch.zeros(2 * K)
        be.call('tss_pwconv_bwd_fused', dict(dz=dz, y=y, lddz=Nc, ldy=Nc, mean=mean, rstd=rstd, gamma=gamma, beta=beta, sums=sums, flags=relu, count=M, dy=dy, lddy=Nc,
                dgamma=dga, dbeta=dbe, wpT=wpT, dx=dx, M=M, K=K, Nc=Nc, lddx=K, yp=yp if link else None, ldyp=K if link else 0, pmean=pm if link else None,
                prstd=pr if link else None, pgamma=pg if link else None, pbeta=pb if link else None, pflags=1 if link else 0, psums=ps if link else None))
        outs[name] = (dy.float(), dx.float(), ps)
    note('pw dy', rel(outs['emu'][0], outs['ref'][0]), 6e-3); note('pw dx', rel(outs['emu'][1], outs['ref'][1]), 1e-2)
    if link: note('pw psums', rel(outs['emu'][2], outs['ref'][2]), 1e-2)
    # ---- dw fused backward
    C = random.choice([32, 64, 96, 128, 192, 384, 576, 768]); N, H, W = random.randint(1, 2), random.randint(1, 19), random.randint(1, 37)
    dt = random.choice([torch.float32, torch.bfloat16]); code = _lib.dtype_code(dt); relu = random.randint(0, 1)
    dz, y, yp = nh(N, C, H, W, dt), nh(N, C, H, W, dt), nh(N, C, H, W, dt)
    w = torch.randn(C, 1, 3, 3, generator=g) / 3
    mean, rstd, gamma, beta = par(C); pm, pr, pg, pb = par(C); M = N * H * W
    sums = torch.zeros(2 * C)
    fake.call('tss_bn_bwd_reduce', dict(dz=dz, z=None, y=y, mean=mean, rstd=rstd, gamma=gamma, beta=beta, sums=sums, M=M, C=C, lddz=C, ldz=0, ldy=C, flags=relu, dtype=code))
    outs = {}
    for name, be in (('ref', fake), ('emu', emu)):
        dy = torch.zeros(N, H, W, C, dtype=dt).permute(0, 3, 1, 2); go = torch.zeros(N, H, W, C, dtype=dt).permute(0, 3, 1, 2)
        dga, dbe, ps = torch.ones(C), torch.ones(C), torch.zeros(2 * C)
        be.call('tss_dwconv3x3_bwd_fused', dict(dz=dz, y=y, w=w, mean=mean, rstd=rstd, gamma=gamma, beta=beta, sums=sums, flags=relu, count=M, dy=dy, dgamma=dga, dbeta=dbe,
                g=go, N=N, H=H, W=W, C=C, yp=yp, pmean=pm, prstd=pr, pgamma=pg, pbeta=pb, pflags=1, psums=ps, dtype=code))
        outs[name] = (dy.float(), go.float(), ps)
    tol = 3e-5 if dt == torch.float32 else 8e-3
    note('dw dy', rel(outs['emu'][0], outs['ref'][0]), tol); note('dw g', rel(outs['emu'][1], outs['ref'][1]), 2 * tol); note('dw psums', rel(outs['emu'][2], outs['ref'][2]), max(tol, 5e-3))
    # ---- stem
    N, H, W = random.randint(1, 2), random.randint(1, 40), random.randint(1, 300)
    x = torch.randn(N, 3, H, W, generator=g); w = torch.randn(32, 3, 3, 3, generator=g) / 5
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    dyy = nh(N, 32, Ho, Wo, torch.bfloat16)
    outs = {}
    for name, be in (('ref', fake), ('emu', emu)):
        yv = torch.zeros(N, Ho, Wo, 32, dtype=torch.bfloat16).permute(0, 3, 1, 2); st = torch.zeros(64, dtype=torch.float64); dw = torch.zeros(32, 3, 3, 3)
        be.call('tss_stem3x3s2_fwd_tc', dict(x=x, w=w, y=yv, N=N, H=H, W=W, Cout=32, scale=None, shift=None, flags=0, stats=st))
        be.call('tss_stem3x3s2_wgrad_tc', dict(x=x, dy=dyy, dw=dw, N=N, H=H, W=W, Cout=32))
        outs[name] = (yv.float(), st, dw)
    note('stem y', rel(outs['emu'][0], outs['ref'][0]), 5e-3); note('stem stats', rel(outs['emu'][1], outs['ref'][1]), 1e-4); note('stem dw', rel(outs['emu'][2], outs['ref'][2]), 5e-3)
    # ---- depthwise fwd / wgrad with the input BatchNorm, pointwise fwd with the input BatchNorm
    C = random.choice([32, 48, 64, 96, 192, 384, 576]); N, H, W = random.randint(1, 2), random.randint(1, 21), random.randint(1, 41)
    stride = random.choice([1, 2]); relu = random.randint(0, 1)
    dt = random.choice([torch.float32, torch.bfloat16]); code = _lib.dtype_code(dt)
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    xx, dyy = nh(N, C, H, W, dt), nh(N, C, Ho, Wo, dt)
    w = torch.randn(C, 1, 3, 3, generator=g) / 3
    sc, sh = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.5 + 0.3
    outs = {}
    for name, be in (('ref', fake), ('emu', emu)):
        yv = torch.zeros(N, Ho, Wo, C, dtype=dt).permute(0, 3, 1, 2); st = torch.zeros(2 * C, dtype=torch.float64); dw = torch.zeros(C, 1, 3, 3)
        be.call('tss_dwconv3x3_fwd_bnin', dict(x=xx, in_scale=sc, in_shift=sh, in_flags=relu, w=w, y=yv, N=N, Hi=H, Wi=W, C=C, stride=stride, stats=st, dtype=code))
        be.call('tss_dwconv3x3_wgrad_bnin', dict(x=xx, in_scale=sc, in_shift=sh, in_flags=relu, dy=dyy, dw=dw, N=N, Hi=H, Wi=W, C=C, stride=stride, dtype=code))
        outs[name] = (yv.float(), st, dw)
    tol = 2e-5 if dt == torch.float32 else 6e-3
    note('dwbnin y', rel(outs['emu'][0], outs['ref'][0]), tol); note('dwbnin stats', rel(outs['emu'][1], outs['ref'][1]), 1e-4); note('dwbnin dw', rel(outs['emu'][2], outs['ref'][2]), 1e-3)
    N, H, W = random.randint(1, 3), random.randint(1, 11), random.randint(1, 13)
    K = 8 * random.randint(1, 96); Nc = 16 * random.randint(1, 8); relu = random.randint(0, 1); M = N * H * W
    xx = nh(N, K, H, W, torch.bfloat16); sc, sh = torch.rand(K, generator=g) + 0.5, torch.randn(K, generator=g) * 0.4
    wp = (torch.randn(Nc, K, generator=g) / K ** 0.5).to(torch.bfloat16)
    outs = {}
    for name, be in (('ref', fake), ('emu', emu)):
        yv = torch.zeros(N, H, W, Nc, dtype=torch.bfloat16).permute(0, 3, 1, 2); zv = torch.zeros(N, H, W, K, dtype=torch.bfloat16).permute(0, 3, 1, 2)
        st = torch.zeros(2 * Nc, dtype=torch.float64)
        be.call('tss_pwconv_fwd_bnin', dict(x=xx, ldx=K, in_scale=sc, in_shift=sh, in_flags=relu, z=zv, ldz=K, wp=wp, y=yv, ldy=Nc, M=M, K=K, Nc=Nc, stats=st))
        outs[name] = (yv.float(), zv.float(), st)
    note('pwbnin y', rel(outs['emu'][0], outs['ref'][0]), 6e-3); note('pwbnin z', rel(outs['emu'][1], outs['ref'][1]), 1e-3); note('pwbnin stats', rel(outs['emu'][2], outs['ref'][2]), 2e-3)
    # ---- stem weight gradient with the BatchNorm-backward apply
    N, H, W = random.randint(1, 2), random.randint(1, 30), random.randint(1, 200); relu = random.randint(0, 1)
    x = torch.randn(N, 3, H, W, generator=g); Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1; M = N * Ho * Wo
    dz, yy = nh(N, 32, Ho, Wo, torch.bfloat16), nh(N, 32, Ho, Wo, torch.bfloat16)
    mean, rstd, gamma, beta = par(32); sums = torch.zeros(64)
    fake.call('tss_bn_bwd_reduce', dict(dz=dz, z=None, y=yy, mean=mean, rstd=rstd, gamma=gamma, beta=beta, sums=sums, M=M, C=32, lddz=32, ldz=0, ldy=32, flags=relu, dtype=1))
    outs = {}
    for name, be in (('ref', fake), ('emu', emu)):
        dw, dga, dbe = torch.zeros(32, 3, 3, 3), torch.zeros(32), torch.zeros(32)
        be.call('tss_stem3x3s2_wgrad_tc_bn', dict(x=x, dz=dz, y=yy, mean=mean, rstd=rstd, gamma=gamma, beta=beta, sums=sums, flags=relu, count=M, dw=dw, dgamma=dga, dbeta=dbe, N=N, H=H, W=W, Cout=32))
        outs[name] = (dw, dga)
    note('stem bn dw', rel(outs['emu'][0], outs['ref'][0]), 6e-3); note('stem bn dgamma', rel(outs['emu'][1], outs['ref'][1]), 1e-6)
    print('trial', trial, 'ok', flush=True)
print({k: '%.2e' % v for k, v in worst.items()})
