"""Where does a CUDA-graph capture of the training step break?  Captures forward + backward + optimizer step of Fast-SCNN
the way engine.GraphedTrainStep does, with the capture status (``tss_capture_status``) queried after every library call
(TSS_CAPTURE_CHECK=1 is forced) and before / after every node of the autograd graph, and prints the first place where the
status turns to "invalidated".  Gates come from the environment (TSS_FUSE_BNIN=1 ...).

    TSS_FUSE_BNIN=1 python tools/debug_capture.py [--batch 2 --height 256 --width 512] [--single-thread]
"""
import argparse
import os
import sys

os.environ['TSS_CAPTURE_CHECK'] = '1'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=2)
    ap.add_argument('--height', type=int, default=256)
    ap.add_argument('--width', type=int, default=512)
    ap.add_argument('--single-thread', action='store_true', help='run the backward pass in the calling thread')
    ap.add_argument('--forward-only', action='store_true')
    ap.add_argument('--unit-grad', action='store_true', help='backward under functional.unit_loss_grad, as the trainer does')
    args = ap.parse_args()
    from torch_semantic_segmentation_b200 import _lib, library
    from torch_semantic_segmentation_b200.functional import wgrad_lane
    from torch_semantic_segmentation_b200.losses import CrossEntropyLoss
    from torch_semantic_segmentation_b200.models import fastscnn
    from torch_semantic_segmentation_b200.optim import FlatAdamW
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    model = fastscnn(3, 19).to(dev).set_compute_dtype(torch.bfloat16).train()
    model.defer_logits = True
    opt = FlatAdamW(model.parameters(), lr=1e-3, weight_decay=1e-5)
    loss_fn = CrossEntropyLoss(ignore_index=255)
    x = torch.randn(args.batch, 3, args.height, args.width, device=dev)
    y = torch.randint(0, 19, (args.batch, args.height, args.width), device=dev)
    if args.single_thread:
        torch.autograd.set_multithreading_enabled(False)

    def status():
        return _lib.backend().call('tss_capture_status', {'stream_handle': torch.cuda.current_stream().cuda_stream})

    log = []
    first_bad = []

    def probe(tag):
        s = status()
        log.append((tag, s, torch.cuda.current_stream().cuda_stream))
        if s not in (0, 1) and not first_bad:
            first_bad.append(len(log) - 1)

    def instrument(root):
        seen, stack = set(), [root]
        while stack:
            node = stack.pop()
            if node is None or node in seen:
                continue
            seen.add(node)
            name = type(node).__name__
            node.register_prehook(lambda grads, name=name: probe('pre  ' + name))
            node.register_hook(lambda gin, gout, name=name: probe('post ' + name))
            for nxt, _ in node.next_functions:
                stack.append(nxt)
        return len(seen)

    def body(instrumented):
        opt.zero_grad()
        loss = loss_fn(model(x), y)
        probe('forward done')
        if args.forward_only:
            return loss
        if instrumented:
            print('autograd nodes: %d' % instrument(loss.grad_fn), flush=True)
        if args.unit_grad:
            from torch_semantic_segmentation_b200.functional import unit_loss_grad
            with unit_loss_grad(loss):
                loss.backward()
        else:
            loss.backward()
        probe('backward done')
        opt.step()
        probe('optimizer done')
        return loss

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            body(False)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    print('eager warm-up ok; lane stream %s, capture stream %#x' % (
        {k: hex(v.cuda_stream) for k, v in wgrad_lane.streams.items()}, side.cuda_stream), flush=True)
    del log[:]
    graph = torch.cuda.CUDAGraph()
    err = None
    try:
        with torch.cuda.graph(graph, stream=side):
            body(True)
    except Exception as exc:        # noqa: BLE001
        err = exc
    if first_bad:
        i = first_bad[0]
        print('capture invalidated between these probes:')
        for tag, s, st in log[max(0, i - 6):i + 3]:
            print('   status %d  stream %#x  %s' % (s, st, tag))
    else:
        print('no probe saw an invalidated capture (%d probes)' % len(log))
    if err is not None:
        print('capture raised: %s: %s' % (type(err).__name__, str(err).splitlines()[0]))
        import traceback
        traceback.print_exception(type(err), err, err.__traceback__, limit=12)
        sys.stdout.flush()
        os._exit(1)
    graph.replay()
    torch.cuda.synchronize()
    print('capture + replay ok')


if __name__ == '__main__':
    main()
