"""``FlatAdamW``: torch.optim.AdamW semantics (scripts/train_fastscnn.py:125-129) over ONE flat
fp32 arena.

On construction every parameter is re-homed into a contiguous fp32 arena (``p.data`` becomes a
view) and gets a persistent gradient view into a second arena (``p.grad``).  Consequences:

* ``zero_grad()`` is one memset, ``step()`` is one streaming kernel (``tss_adamw_step``);
* the weight-gradient kernels accumulate straight into the arena (``p._tss_grad``), so
  autograd launches no per-parameter accumulation kernels;
* data-parallel training all-reduces contiguous slices of the gradient arena (buckets)
  without flatten / unflatten copies (``distributed.GradientAllReducer``).

Hyper-parameters live in a small device tensor refreshed from pinned host memory each step,
so a captured CUDA graph of the step can be replayed under a learning-rate schedule.
"""
import torch

from . import ops

__all__ = ['FlatAdamW']


class FlatAdamW(torch.optim.Optimizer):

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        if len(self.param_groups) != 1:
            raise ValueError('FlatAdamW supports a single parameter group')
        ps = [p for p in self.param_groups[0]['params'] if p.requires_grad]
        if not ps:
            raise ValueError('no trainable parameters')
        dev = ps[0].device
        if any(p.device != dev or p.dtype != torch.float32 for p in ps):
            raise ValueError('FlatAdamW needs fp32 parameters on one device')
        # 16-byte aligned slots so that every view is vector-load friendly
        offsets, n = [], 0
        for p in ps:
            offsets.append(n)
            n += (p.numel() + 3) // 4 * 4
        self.numel = n
        self.param_arena = torch.zeros(n, dtype=torch.float32, device=dev)
        self.grad_arena = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self.slots = []
        with torch.no_grad():
            for p, off in zip(ps, offsets):
                view = self.param_arena[off:off + p.numel()].view_as(p)
                view.copy_(p)
                p.data = view
                g = self.grad_arena[off:off + p.numel()].view_as(p)
                p.grad = g
                p._tss_grad = g          # kernels accumulate here directly (functional.py)
                p._tss_optimizer = self  # pointwise blocks register their bf16 packs here (register_pack)
                self.slots.append((p, off, p.numel()))
        g0 = self.param_groups[0]
        self.hyper = torch.tensor([g0['lr'], g0['betas'][0], g0['betas'][1], g0['eps'], g0['weight_decay'],
                                   0.0, 0.0, 0.0], dtype=torch.float32, device=dev)
        # two pinned staging buffers used in turn: the host never rewrites the one an in-flight copy may still read
        self._hosts = [torch.empty(5, dtype=torch.float32) for _ in range(2)]
        if dev.type == 'cuda':
            self._hosts = [h.pin_memory() for h in self._hosts]
        self._host = self._hosts[0]
        self._host_turn = 0
        self.grad_scale = 1.0          # e.g. 1/world_size after a SUM all-reduce
        self._offsets = {id(p): off for p, off, _ in self.slots}
        self._packs = {}               # id(param) -> (wp, wpT): bf16 TMA/UMMA operands of the pointwise convs
        self._pack_rows = []
        self._pack_table = None
        self._pack_max = 0

    def register_pack(self, p):
        """Persistent bf16 copies (W, W^T) of a pointwise weight that lives in the arena.  They are packed
        now and then refreshed by ONE multi-tensor launch after every ``step()`` (instead of one launch per
        layer per forward), at fixed addresses (CUDA-graph friendly)."""
        hit = self._packs.get(id(p))
        if hit is None:
            Nc, K = p.shape[0], p[0].numel()
            wp = torch.empty((Nc, K), dtype=torch.bfloat16, device=p.device)
            wpT = torch.empty((K, Nc), dtype=torch.bfloat16, device=p.device)
            hit = self._packs[id(p)] = (wp, wpT)
            self._pack_rows.append([self._offsets[id(p)], Nc, K, wp.data_ptr(), wpT.data_ptr()])
            self._pack_table = torch.tensor(self._pack_rows, dtype=torch.int64, device=p.device)
            self._pack_max = max(self._pack_max, Nc * K)
        from . import _lib
        _lib.call('tss_pack_weights_bf16', w=p.detach(), wp=hit[0], wpT=hit[1], Nc=hit[0].shape[0], K=hit[0].shape[1])
        return hit

    def _repack(self):
        if self._pack_table is not None:
            from . import _lib
            _lib.call('tss_pack_weights_multi', arena=self.param_arena, table=self._pack_table,
                      n_entries=len(self._pack_rows), max_elems=self._pack_max)

    def zero_grad(self, set_to_none=False):
        """One memset; the gradient views stay attached (``set_to_none`` is ignored)."""
        from .functional import wgrad_lane
        wgrad_lane.join(self.grad_arena.device if self.grad_arena.is_cuda else None)
        self.grad_arena.zero_()

    def _fill(self, buf):
        g = self.param_groups[0]
        buf[0], buf[1], buf[2] = g['lr'], g['betas'][0], g['betas'][1]
        buf[3], buf[4] = g['eps'], g['weight_decay']

    def write_host_hyper(self):
        """param_groups -> the pinned staging buffers of the hyper-parameters.  Host work only.  A captured
        ``step()`` replays the pinned->device copy from whichever buffer was current at capture time, so a CUDA-graph
        replay picks up whatever this wrote last: call it before every replay (engine.GraphedTrainStep does) and
        learning-rate schedules keep working under graphs.  (With ``lazy_loss`` the host runs one step ahead of the device:
        the copy at the end of step i may then already see the values written for step i+1.)"""
        for buf in self._hosts:
            self._fill(buf)

    def _refresh_hyper(self):
        capturing = self.hyper.is_cuda and torch.cuda.is_current_stream_capturing()
        if not capturing:
            # eager mode: the two staging buffers take turns, so that the host never rewrites the one an in-flight
            # asynchronous copy of the previous step may still be reading
            self._host_turn ^= 1
            self._host = self._hosts[self._host_turn]
        self._fill(self._host)
        self.hyper[:5].copy_(self._host, non_blocking=True)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        from .functional import wgrad_lane
        wgrad_lane.join(self.grad_arena.device if self.grad_arena.is_cuda else None)   # side-stream wgrads
        reducer = getattr(self, '_reducer', None)
        if reducer is not None:
            reducer.finish()         # gradient all-reduce buckets launched during backward
        self._refresh_hyper()
        ops.adamw_step(self.param_arena, self.grad_arena, self.exp_avg, self.exp_avg_sq, self.hyper,
                       self.grad_scale)
        self._repack()
        return loss

    @property
    def step_count(self):
        return int(self.hyper[5].item())

    # torch.optim.Optimizer keeps per-parameter state in ``self.state``; here the moments are two flat arenas and the
    # step counter lives on the device, so checkpointing (ModelCheckpoint in the reference's scripts saves
    # ``optimizer.state_dict()``) has to carry them explicitly.
    def state_dict(self):
        sd = super().state_dict()
        sd['flat_adamw'] = {'exp_avg': self.exp_avg.detach().cpu().clone(), 'exp_avg_sq': self.exp_avg_sq.detach().cpu().clone(),
                            'counters': self.hyper[5:].detach().cpu().clone(), 'numel': self.numel}
        return sd

    def load_state_dict(self, state_dict):
        state_dict = dict(state_dict)
        flat = state_dict.pop('flat_adamw', None)
        super().load_state_dict(state_dict)
        if flat is not None:
            if int(flat['numel']) != self.numel:
                raise ValueError('FlatAdamW.load_state_dict: arena of %d elements, checkpoint has %d' % (self.numel, int(flat['numel'])))
            self.exp_avg.copy_(flat['exp_avg'])
            self.exp_avg_sq.copy_(flat['exp_avg_sq'])
            self.hyper[5:].copy_(flat['counters'])
        self.write_host_hyper()
