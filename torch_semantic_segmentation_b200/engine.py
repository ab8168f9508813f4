"""Train / eval loop, drop-in for ``torch_semantic_segmentation.engine`` (reference: engine.py).

``create_segmentation_trainer`` and ``create_segmentation_evaluator`` keep the reference's
signatures and step semantics (engine.py:22-39, 59-82).  The reference builds them on
pytorch-ignite and NVIDIA apex, neither of which this path needs:

* the small :class:`Engine` / :class:`Events` below provide the part of the ignite API the
  reference scripts use (``run``, ``state``, ``on`` / ``add_event_handler``, ``every=``
  filters, metric dict);
* ``use_f16=True`` selects the bf16 kernels (fp32 master weights, fp32 BatchNorm statistics;
  bf16 has fp32's exponent range, so apex's loss scaling, engine.py:32-34, has no counterpart);
* the evaluator feeds ONE device-side confusion matrix (the reference keeps four identical
  ones, engine.py:65-72) and derives ``iou`` / ``miou`` / ``accuracy`` / ``dice`` from it.
"""
import os

import torch

from . import metrics as M
from . import functional as Fn
from . import library
from .gates import gate
from .functional import unit_loss_grad

__all__ = ['create_segmentation_trainer', 'create_segmentation_evaluator', 'Engine', 'Events', 'State']


# ------------------------------------------------------------------ minimal ignite surface --
class _Event:
    def __init__(self, name, every=None, once=None):
        self.name, self.every, self.once = name, every, once

    def __call__(self, every=None, once=None):
        return _Event(self.name, every, once)

    def __eq__(self, other):
        return isinstance(other, _Event) and other.name == self.name

    def __hash__(self):
        return hash(self.name)

    def __repr__(self):
        return 'Events.%s' % self.name


class Events:
    STARTED = _Event('STARTED')
    EPOCH_STARTED = _Event('EPOCH_STARTED')
    ITERATION_STARTED = _Event('ITERATION_STARTED')
    ITERATION_COMPLETED = _Event('ITERATION_COMPLETED')
    EPOCH_COMPLETED = _Event('EPOCH_COMPLETED')
    COMPLETED = _Event('COMPLETED')


_END = object()


class State:
    def __init__(self):
        self.iteration = 0
        self.epoch = 0
        self.max_epochs = None
        self.output = None
        self.batch = None
        self.metrics = {}


class Engine:
    """``Engine(process_function)``; ``process_function(engine, batch) -> output``."""

    def __init__(self, process_function):
        self._process_function = process_function
        self._handlers = {}
        self.state = State()
        self.should_terminate = False
        # optional ``stage(batch) -> staged``: called for batch i+1 BEFORE batch i is processed, so that
        # its host->device copy (side stream) overlaps with the step running on batch i
        self.stage = None
        self._pending, self._staged_next = _END, None

    def add_event_handler(self, event, handler, *args, **kwargs):
        self._handlers.setdefault(event.name, []).append((event, handler, args, kwargs))

    def on(self, event, *args, **kwargs):
        def decorator(fn):
            self.add_event_handler(event, fn, *args, **kwargs)
            return fn
        return decorator

    def terminate(self):
        self.should_terminate = True

    def prefetch_next(self):
        """Issue the host->device copy of the next batch now (idempotent; a no-op without ``stage``)."""
        nxt = getattr(self, '_pending', _END)
        if self.stage is not None and nxt is not _END and self._staged_next is None:
            self._staged_next = self.stage(nxt)

    def _fire(self, event):
        count = self.state.epoch if 'EPOCH' in event.name else self.state.iteration
        for ev, handler, args, kwargs in self._handlers.get(event.name, ()):
            if ev.every is not None and (count == 0 or count % ev.every != 0):
                continue
            if ev.once is not None and count != ev.once:
                continue
            handler(self, *args, **kwargs)

    def run(self, data, max_epochs=1):
        self.state = State()
        self.state.max_epochs = max_epochs
        self.should_terminate = False
        self._fire(Events.STARTED)
        for _ in range(max_epochs):
            if self.should_terminate:
                break
            self.state.epoch += 1
            self._fire(Events.EPOCH_STARTED)
            it = iter(data)
            nxt = next(it, _END)
            self._pending, self._staged_next = nxt, None
            self.prefetch_next()
            while nxt is not _END:
                batch, current = nxt, self._staged_next
                nxt = next(it, _END)
                self._pending, self._staged_next = nxt, None
                self.state.iteration += 1
                self.state.batch = batch
                self._fire(Events.ITERATION_STARTED)
                # the process function may call prefetch_next() right after it has LAUNCHED its step (and
                # before it waits for the result); otherwise the next batch is staged when it returns
                self.state.output = self._process_function(self, batch if current is None else current)
                self.prefetch_next()
                self._fire(Events.ITERATION_COMPLETED)
                if self.should_terminate:
                    break
            self._fire(Events.EPOCH_COMPLETED)
        self._fire(Events.COMPLETED)
        return self.state


class RunningAverage:
    """Exponential moving average of the step output (ignite default alpha = 0.98)."""

    def __init__(self, output_transform=lambda x: x, alpha=0.98):
        self.transform, self.alpha, self.value = output_transform, alpha, None

    def attach(self, engine, name):
        def started(_engine):
            self.value = None

        def update(e):
            v = float(self.transform(e.state.output))
            self.value = v if self.value is None else self.value * self.alpha + (1.0 - self.alpha) * v
            e.state.metrics[name] = self.value
        engine.add_event_handler(Events.EPOCH_STARTED, started)
        engine.add_event_handler(Events.ITERATION_COMPLETED, update)


def _prepare_batch(batch, device=None, non_blocking=False):
    """ignite's ``_prepare_batch`` (engine.py:27): move (x, y) to the device."""
    x, y = batch
    return (x.to(device=device, non_blocking=non_blocking),
            y.to(device=device, non_blocking=non_blocking))


# One CUDA graph per staging slot of the trainer (no device-to-device copy of the batch into the graph's inputs).
# Host-side change only; on by default (validated on the B200: tests/test_fused_paths_gpu.py).
SLOT_GRAPHS = gate('SLOT_GRAPHS')


class _StagedBatch:
    """A batch whose host->device copy was issued ahead of time on a side stream."""

    __slots__ = ('x', 'y', 'ready', 'slot')

    def __init__(self, x, y, ready, slot):
        self.x, self.y, self.ready, self.slot = x, y, ready, slot


class _BatchStager:
    """Double-buffered host->device staging: ``stage(batch)`` copies (x, y) into one of two device
    slots on a copy stream and returns at once; ``take(staged)`` makes the compute stream wait for
    that copy.  A slot is overwritten only after the step that consumed it has been enqueued
    (``release``), which the copy stream waits for."""

    def __init__(self, device, non_blocking=True):
        self.device = torch.device(device)
        self.non_blocking = non_blocking
        self.stream = torch.cuda.Stream(self.device)
        self.slots = [None, None]
        self.free = [None, None]        # event: the consumer of the slot is done with it
        self.turn = 0

    def stage(self, batch):
        x, y = batch
        k = self.turn
        self.turn ^= 1
        slot = self.slots[k]
        if slot is None or slot[0].shape != x.shape or slot[0].dtype != x.dtype or slot[1].shape != y.shape \
                or slot[1].dtype != y.dtype:
            slot = self.slots[k] = (torch.empty(x.shape, dtype=x.dtype, device=self.device),
                                    torch.empty(y.shape, dtype=y.dtype, device=self.device))
            self.free[k] = None
            # the caching allocator may have handed out a block that kernels already enqueued on the compute stream
            # still use (freed on the host, not yet on the device): the first copy into a NEW slot waits for them
            self.stream.wait_stream(torch.cuda.current_stream(self.device))
        if self.free[k] is not None:
            self.stream.wait_event(self.free[k])
        with torch.cuda.stream(self.stream):
            slot[0].copy_(x, non_blocking=self.non_blocking)
            slot[1].copy_(y, non_blocking=self.non_blocking)
            ready = torch.cuda.Event()
            ready.record(self.stream)
        return _StagedBatch(slot[0], slot[1], ready, k)

    def take(self, staged):
        torch.cuda.current_stream(self.device).wait_event(staged.ready)
        return staged.x, staged.y

    def release(self, staged):
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.free[staged.slot] = ev


class _LossReader:
    """What ``update_fn`` returns for the step's loss.  ``lazy=False``: ``loss.item()`` -- the reference's contract
    (engine.py:39), one host synchronisation per step, during which the GPU idles while the host prepares the next
    launch.  ``lazy=True``: the loss is copied to pinned host memory behind the step and the PREVIOUS step's value is
    returned (the first step returns its own), so the host always runs one step ahead of the device."""

    def __init__(self, device, lazy):
        self.lazy = lazy and torch.device(device).type == 'cuda'
        self.slots = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(2)] if self.lazy else None
        self.k, self.pending = 0, None

    def __call__(self, loss):
        if not self.lazy:
            return loss.item()
        k, self.k = self.k, self.k ^ 1
        self.slots[k].copy_(loss.detach().reshape(1).float(), non_blocking=True)
        done = torch.cuda.Event()
        done.record()
        prev, self.pending = self.pending, (k, done)
        if prev is None:
            prev = self.pending
        prev[1].synchronize()
        return float(self.slots[prev[0]])


_capture_streams = {}


def _capture_stream(dev):
    """ONE warm-up / capture stream per device for every graphed step of the process.  autograd remembers the stream on
    which a parameter's AccumulateGrad node was created and makes the end of every backward pass that runs the node wait
    for that stream; a node that outlives its step (any lingering reference to a tensor with a ``grad_fn``) would carry a
    foreign, non-capturing stream into the next capture ("dependency created on uncaptured work in another stream")."""
    dev = torch.device(dev)
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    st = _capture_streams.get(key)
    if st is None:
        # TSS_MAIN_PRIORITY=1 gives the stream of the step's critical chain a high priority (captured kernel nodes inherit it)
        # over the weight-gradient lane and the all-reduce stream.  Measured on B200: 3.51 against 3.43 ms/step -- the side
        # kernels are pushed behind the chain and leave a longer tail than the dispatch stalls they cause; off by default.
        prio = -1 if os.environ.get('TSS_MAIN_PRIORITY', '0') == '1' else 0
        st = _capture_streams[key] = torch.cuda.Stream(dev, priority=prio)
    return st


class GraphedTrainStep:
    """The whole optimisation step (zero_grad, forward, loss, backward, gradient all-reduce,
    optimizer) captured ONCE as a CUDA graph and replayed per batch: ~350 kernel launches become
    one graph launch, which is what a 1.1 M-parameter net needs to stay GPU-bound.  The batch is
    copied straight from (pinned) host memory into the graph's static input buffers."""

    def __init__(self, model, optimizer, loss_fn, x, y, warmup=3, transform=None, bind=False):
        dev = x.device
        if bind:      # the graph reads (x, y) in place: a staging slot of the trainer, refilled by its copy stream
            self.x, self.y = x, y
        else:
            self.x = torch.empty_like(x)
            self.y = torch.empty_like(y)
            self.x.copy_(x)
            self.y.copy_(y)
        self.key = (tuple(x.shape), x.dtype, tuple(y.shape), y.dtype)
        # device-side input pipeline (data.DeviceTransform): the graph's inputs are the decoded uint8 frames; the
        # random draws of a step live in a small static table that is refreshed before every replay
        self.optimizer = optimizer
        self.transform = transform
        self.geom = None
        if transform is not None:
            self.geom = transform.draw_geometry(x.shape[0], x.shape[1], x.shape[2]).to(dev)

        def body():
            optimizer.zero_grad()
            xs, ys = (self.x, self.y) if transform is None else transform.apply(self.x, self.y, self.geom)
            loss = loss_fn(model(xs), ys)
            if library._CAPTURE_CHECK and torch.cuda.is_current_stream_capturing():
                library.instrument_backward(loss.grad_fn)
            with unit_loss_grad(loss):
                loss.backward()
            if library._CAPTURE_CHECK:
                library.check_capture('the end of backward')
            optimizer.step()
            return loss.detach()

        side = _capture_stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        self.warmup_loss = None
        with torch.cuda.stream(side):
            for _ in range(warmup):          # allocator / cuBLAS-free warm-up on a side stream
                self.warmup_loss = body()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        from . import _lib
        self.graph = torch.cuda.CUDAGraph()
        before = _lib.launch_count()
        # capture on the SAME side stream the warm-up ran on: autograd remembers the stream on which a parameter's
        # AccumulateGrad first ran and makes the end of every later backward wait for it -- a foreign, non-capturing
        # stream at that point invalidates the capture ("dependency created on uncaptured work in another stream")
        with torch.cuda.graph(self.graph, stream=side):
            self.loss = body()
        self.kernels_per_step = _lib.launch_count() - before   # libtss_b200 kernel nodes in the graph
        self.warmup_steps = warmup            # steps the optimizer really took on the example batch (capture runs nothing)

    def matches(self, x, y):
        return (tuple(x.shape), x.dtype, tuple(y.shape), y.dtype) == self.key

    def __call__(self, x, y, non_blocking=True):
        if x is not self.x:
            self.x.copy_(x, non_blocking=non_blocking)
        if y is not self.y:
            self.y.copy_(y, non_blocking=non_blocking)
        if self.transform is not None:
            self.geom.copy_(self.transform.draw_geometry(x.shape[0], x.shape[1], x.shape[2]), non_blocking=non_blocking)
        refresh = getattr(self.optimizer, 'write_host_hyper', None)
        if refresh is not None:
            refresh()        # a learning-rate scheduler may have changed param_groups since the last replay
        self.graph.replay()
        # the replay rewrote parameters and BatchNorm buffers behind autograd's back, without running the Python that
        # bumps this counter in eager mode: caches derived from them (eval-mode folded BatchNorm, weight packs) are stale
        from . import ops
        ops.WEIGHTS_EPOCH[0] += 1
        return self.loss


class GraphedEvalStep:
    """Inference forward + confusion-matrix update captured as ONE CUDA graph per batch shape (the eval
    forward is ~55 kernels of 5-20 us: launched eagerly from Python it is host-bound at small batches)."""

    def __init__(self, model, cm, x, y, warmup=2):
        dev = x.device
        self.x, self.y = torch.empty_like(x), torch.empty_like(y)
        self.x.copy_(x)
        self.y.copy_(y)
        self.key = (tuple(x.shape), x.dtype, tuple(y.shape), y.dtype)
        cm._ensure(dev)
        keep = cm.cm.clone()

        def body():
            with torch.no_grad():
                out = model(self.x)
                cm.update((out, self.y))
            return out

        side = _capture_stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                body()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=side):
            self.out = body()
        torch.cuda.synchronize(dev)
        cm.cm.copy_(keep)                 # warm-up and capture did not count
        cm.num_examples -= (warmup + 1) * y.shape[0]

    def matches(self, x, y):
        return (tuple(x.shape), x.dtype, tuple(y.shape), y.dtype) == self.key

    def __call__(self, x, y, non_blocking=True):
        self.x.copy_(x, non_blocking=non_blocking)
        self.y.copy_(y, non_blocking=non_blocking)
        self.graph.replay()
        return self.out, self.y


# ------------------------------------------------------------------ the two factories -------
def create_segmentation_trainer(model, optimizer, loss_fn, device, use_f16=False, logging=True,
                                non_blocking=True, cuda_graph=False, prefetch=True, transform=None, lazy_loss=False):
    """reference: engine.py:22-56.  ``cuda_graph=True`` (an addition) replays the step as one
    CUDA graph; it needs fixed batch shapes and a capturable optimizer (``optim.FlatAdamW`` or
    ``torch.optim.AdamW(capturable=True)``).  The first batch is used for warm-up and capture
    (its optimisation steps are real steps).  ``prefetch=True`` (an addition) issues the host->device
    copy of batch i+1 on a copy stream before the step on batch i is launched, so the copy (141 MB per
    step of 12 x 768 x 768 fp32 images + int64 labels) overlaps with compute; ``engine.py:27`` copies
    synchronously in front of every step.  ``transform`` (an addition): a ``data.DeviceTransform``; the loader then
    yields decoded uint8 frames and label ids ((N,H,W,3), (N,H,W)) and scale / crop / flip / normalise / label
    mapping run on the device in front of the model (scripts/train_fastscnn.py:62-72 does them in the workers).
    ``lazy_loss=True`` (an addition): ``state.output`` of iteration i is the loss of iteration i-1 (``_LossReader``), which
    removes the per-step host synchronisation of ``engine.py:39`` from the critical path."""
    if use_f16 and hasattr(model, 'set_compute_dtype'):
        model.set_compute_dtype(torch.bfloat16)
    Fn.enable_deferred_logits(model, loss_fn)     # nobody but loss_fn sees y_pred inside update_fn
    graphed = {}
    stager = _BatchStager(device, non_blocking) if (prefetch and torch.device(device).type == 'cuda') else None
    read_loss = _LossReader(device, lazy_loss)

    def fetch(batch):
        """-> (x, y) on the device; a staged batch only has to be waited for."""
        if isinstance(batch, _StagedBatch):
            return stager.take(batch)
        return _prepare_batch(batch, device=device, non_blocking=non_blocking)

    def update_fn(_trainer, batch):
        model.train()
        staged = batch if isinstance(batch, _StagedBatch) else None
        if cuda_graph:
            # SLOT_GRAPHS (gated): one captured graph per staging slot, reading the slot in place -- no device-to-device
            # copy of the batch into the graph's inputs (141 MB read + written per step at the benchmark shape)
            key = ('step', staged.slot) if (SLOT_GRAPHS and staged is not None) else 'step'
            g = graphed.get(key)
            if g is None:
                xd, yd = fetch(batch)
                first = not graphed
                graphed[key] = g = GraphedTrainStep(model, optimizer, loss_fn, xd, yd, transform=transform,
                                                    warmup=3 if first else 0, bind=key != 'step')
                if first:
                    # capture executes nothing and g.loss is a never-written buffer of the graph's pool: the step(s) on
                    # this batch were the warm-up ones, its loss is the last warm-up step's
                    out = g.warmup_loss
                else:
                    g.graph.replay()         # this is the step on the current batch
                    out = g.loss
                if staged is not None:
                    stager.release(staged)
                if SLOT_GRAPHS:
                    _trainer.prefetch_next()
                return read_loss(out)
            x, y = fetch(batch) if staged is not None else batch
            if not g.matches(x, y):
                raise RuntimeError('cuda_graph=True needs a fixed batch shape; got %s' % (tuple(x.shape),))
            loss = g(x, y, non_blocking)         # device->device into the graph's inputs when staged
            if staged is not None:
                stager.release(staged)
            _trainer.prefetch_next()             # the step is in flight: now copy batch i+1 under it
            return read_loss(loss)
        optimizer.zero_grad()
        x, y = fetch(batch)
        if transform is not None:
            x, y = transform(x, y)

        y_pred = model(x)
        loss = loss_fn(y_pred, y)

        with unit_loss_grad(loss):      # backward() seeds d(loss)/d(loss) = 1: no rescale pass needed
            loss.backward()

        optimizer.step()
        if staged is not None:
            stager.release(staged)
        _trainer.prefetch_next()
        return read_loss(loss)

    trainer = Engine(update_fn)
    if stager is not None:
        trainer.stage = stager.stage
    RunningAverage(output_transform=lambda x: x).attach(trainer, 'loss')

    @trainer.on(Events.ITERATION_COMPLETED)
    def log_optimizer_params(engine):
        param_groups = optimizer.param_groups[0]
        for h in ['lr', 'momentum', 'weight_decay']:
            if h in param_groups.keys():
                engine.state.metrics[h] = param_groups[h]

    if logging:
        @trainer.on(Events.EPOCH_COMPLETED)
        def _log(engine):
            print('epoch %d iteration %d loss %.4f lr %s' % (
                engine.state.epoch, engine.state.iteration, engine.state.metrics.get('loss', float('nan')),
                engine.state.metrics.get('lr')))

    return trainer


def create_segmentation_evaluator(model, device, num_classes=19, loss_fn=None, non_blocking=True,
                                  cuda_graph=False):
    """reference: engine.py:59-82.  ``state.metrics`` gets 'iou', 'miou', 'accuracy', 'dice'
    (float64 tensors / scalars, ignite formulas) and, with ``loss_fn``, 'loss'.  ``cuda_graph=True`` (an
    addition, without ``loss_fn``) replays forward + confusion-matrix update as one CUDA graph per batch
    shape; the returned ``y_pred`` is then the graph's output buffer (overwritten by the next batch)."""
    cm = M.ConfusionMatrix(num_classes)
    loss_acc = {'sum': None, 'n': 0}
    graphs = {}

    def eval_fn(_evaluator, batch):
        model.eval()
        if cuda_graph and loss_fn is None:
            x, y = batch
            key = (tuple(x.shape), x.dtype, tuple(y.shape), y.dtype)
            g = graphs.get(key)
            if g is None:
                xd, yd = _prepare_batch(batch, device=device, non_blocking=non_blocking)
                g = graphs[key] = GraphedEvalStep(model, cm, xd, yd)
            out = g(x, y, non_blocking)
            cm.num_examples += y.shape[0]
            return out
        with torch.no_grad():
            x, y = _prepare_batch(batch, device=device, non_blocking=non_blocking)
            y_pred = model(x)
            cm.update((y_pred, y))
            if loss_fn is not None:
                l = loss_fn(y_pred, y).detach().double() * y.shape[0]
                loss_acc['sum'] = l if loss_acc['sum'] is None else loss_acc['sum'] + l
                loss_acc['n'] += y.shape[0]
            return y_pred, y

    evaluator = Engine(eval_fn)

    @evaluator.on(Events.EPOCH_STARTED)
    def _reset(engine):
        cm.reset()
        loss_acc['sum'], loss_acc['n'] = None, 0
        if graphs:
            # captured forwards read the folded BatchNorm / packed weights by address: refresh them in place first
            from .nn.blocks import refresh_cached_operands
            model.eval()
            refresh_cached_operands(model)

    @evaluator.on(Events.EPOCH_COMPLETED)
    def _compute(engine):
        engine.state.metrics.update(M.metrics_from_cm(cm.compute()))
        engine.state.metrics['confusion_matrix'] = cm.compute(sync=False)
        if loss_fn is not None and loss_acc['n'] > 0:
            engine.state.metrics['loss'] = float(loss_acc['sum']) / loss_acc['n']

    evaluator.confusion_matrix = cm
    return evaluator
