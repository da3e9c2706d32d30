"""The loss path inside CUDA graphs.  The library allocates nothing, keeps no global state and never
synchronises the host, so `criterion(preds, annotations)` -- no-grad forward as well as the training
forward + backward -- can be captured with torch.cuda.graph and replayed; the small-batch configs
(BASELINE configs 1-3) are launch- / host-bound in eager mode (tools/prof_small.py times both).
The sums are 64-bit fixed point (order-independent), so a replay is bit-identical to an eager call."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(kind, batch=2, size=256, classes=8):
    from b200det import synth, losses
    if kind == 'retina':
        preds = synth.make_retina_preds(batch, size, classes, seed=4, device='cuda')
        crit = losses.RetinaLoss(**synth.RETINA_KW, box_loss_type='GIoU')
    else:
        preds = synth.make_fcos_preds(batch, size, classes, seed=4, device='cuda')
        crit = losses.FCOSLoss(strides=synth.STRIDES, mi=synth.MI)
    ann = synth.make_annotations(batch, 20, size, classes, seed=5).cuda()
    return preds, ann, crit


def _warm(fn):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()


@pytest.mark.parametrize('kind', ['retina', 'fcos'])
def test_no_grad_loss_replays_in_a_cuda_graph(kind):
    preds, ann, crit = _setup(kind)

    def fwd():
        with torch.no_grad():
            return crit(preds, ann)

    _warm(fwd)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        static = fwd()
    for round_ in range(2):
        for t in preds[0]:
            t.mul_(0.9)                     # new head outputs in the same buffers
        ann[0, 0, :4] += 1.0
        graph.replay()
        got = {k: v.clone() for k, v in static.items()}
        want = fwd()
        for k in want:
            assert torch.equal(got[k], want[k]), (kind, k, round_)
            assert torch.isfinite(got[k]).all() and got[k].item() > 0


@pytest.mark.parametrize('kind', ['retina', 'fcos'])
def test_training_step_replays_in_a_cuda_graph(kind):
    preds, ann, crit = _setup(kind)
    req = [[t.clone().requires_grad_(True) for t in grp] for grp in preds]
    flat = [t for grp in req for t in grp]

    def step():
        d = crit(req, ann)
        sum(d.values()).backward()
        return d

    _warm(step)
    for t in flat:
        t.grad = None
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        static = step()
    with torch.no_grad():
        for t in req[0]:
            t.mul_(0.9)
    graph.replay()
    torch.cuda.synchronize()
    got_loss = {k: v.detach().clone() for k, v in static.items()}
    got_grad = [t.grad.clone() for t in flat]
    for t in flat:
        t.grad = None
    want = step()
    for k in want:
        assert torch.equal(got_loss[k], want[k].detach()), (kind, k)
    for g, t in zip(got_grad, flat):
        assert torch.equal(g, t.grad), kind
    assert any(float(g.abs().sum()) > 0 for g in got_grad)
