"""Drop-in check from the reference's REAL config files: tests/golden/config_kwargs.json holds the
kwargs dict of every `losses.__dict__[name](**{...})` / `decode.__dict__[name](**{...})` call in
3.detection_training/**/{train,test}_config.py and 10.face_detection_training/**/*_config.py (60
calls, extracted with ast by tests/golden/make_config_kwargs.py).  The b200det modules must build
every one of them by the same lookup, with the same attribute values as the reference's classes
(CPU), and run them (GPU)."""
import json
import os

import numpy as np
import pytest
import torch

import b200det
from b200det import losses, decode, synth
from b200det.face_detection import losses as face_losses, decode as face_decode

import refload

HERE = os.path.dirname(os.path.abspath(__file__))
ENTRIES = json.load(open(os.path.join(HERE, 'golden', 'config_kwargs.json')))


def module_for(entry):
    face = entry['class'].startswith('RetinaFace')
    if entry['module'] == 'losses':
        return face_losses if face else losses
    return face_decode if face else decode


def distinct():
    seen, out = set(), []
    for e in ENTRIES:
        key = (e['class'], json.dumps(e['kwargs'], sort_keys=True))
        if key not in seen:
            seen.add(key)
            out.append(e)
    return out


def test_every_reference_config_call_constructs():
    assert len(ENTRIES) >= 60
    classes = {e['class'] for e in ENTRIES}
    assert {'RetinaLoss', 'FCOSLoss', 'RetinaDecoder', 'FCOSDecoder', 'RetinaFaceLoss',
            'RetinaFaceDecoder', 'DETRDecoder', 'DINODETRDecoder'} <= classes
    for e in ENTRIES:
        mod = module_for(e)
        assert e['class'] in mod.__dict__, f"{e['class']} missing from {mod.__name__}.__dict__"
        obj = mod.__dict__[e['class']](**e['kwargs'])     # the config's own call
        for k, v in e['kwargs'].items():
            assert getattr(obj, k) == v, f"{e['file']}:{e['line']} {e['class']}.{k}"
        if e['module'] == 'losses':
            assert isinstance(obj, torch.nn.Module) and callable(getattr(obj, 'forward'))
        else:
            assert callable(obj)


@pytest.mark.skipif(not refload.available(), reason='reference checkout not present')
def test_attributes_equal_the_reference_objects():
    """The same kwargs through the reference's own classes give objects with the same public
    attribute values (what the training / test scripts and checkpoints may read)."""
    L, D, _ = refload.load()
    FL, FD = refload.load_face()
    for e in distinct():
        face = e['class'].startswith('RetinaFace')
        ref_mod = (FL if face else L) if e['module'] == 'losses' else (FD if face else D)
        theirs = ref_mod.__dict__[e['class']](**e['kwargs'])
        ours = module_for(e).__dict__[e['class']](**e['kwargs'])
        for k, v in vars(theirs).items():
            if k.startswith('_') or isinstance(v, (torch.nn.Module,)) or callable(v):
                continue
            if k in ('anchors', 'positions', 'decode_function', 'iou_function', 'training'):
                continue
            assert hasattr(ours, k), f"{e['class']}: attribute {k} missing"
            assert getattr(ours, k) == v, f"{e['class']}.{k}: {getattr(ours, k)} != {v}"


@pytest.mark.gpu
@pytest.mark.parametrize('entry', distinct(), ids=lambda e: e['class'])
def test_config_objects_run_on_the_gpu(entry):
    """One tiny forward per distinct kwargs set, checked against the oracle."""
    from oracle import det_oracle as O
    obj = module_for(entry).__dict__[entry['class']](**entry['kwargs'])
    name, kw = entry['class'], entry['kwargs']
    if name in ('RetinaLoss', 'RetinaDecoder'):
        preds = synth.make_tie_free(synth.make_retina_preds(2, 128, 8, seed=80))
        ann = synth.make_annotations(2, 12, 128, 8, seed=81)
        dpreds = synth.to_device(preds, 'cuda')
        if name == 'RetinaLoss':
            with torch.no_grad():
                got = obj.cuda()(dpreds, ann.cuda())
                want = O.retina_loss(preds, ann, **kw)
            for k in ('cls_loss', 'reg_loss'):
                np.testing.assert_allclose(got[k].item(), want[k].item(), rtol=1e-5)
        else:
            got = obj(dpreds)
            want, _ = O.retina_decode(preds, **kw)
            for a, b in zip(got, want):
                assert np.array_equal(a, b)
    elif name in ('FCOSLoss', 'FCOSDecoder'):
        preds = synth.make_tie_free(synth.make_fcos_preds(2, 256, 8, seed=82))
        ann = synth.make_annotations(2, 12, 256, 8, seed=83)
        dpreds = synth.to_device(preds, 'cuda')
        if name == 'FCOSLoss':
            with torch.no_grad():
                got = obj.cuda()(dpreds, ann.cuda())
                want = O.fcos_loss(preds, ann, **kw)
            for k in ('cls_loss', 'reg_loss', 'center_ness_loss'):
                np.testing.assert_allclose(got[k].item(), want[k].item(), rtol=1e-5)
        else:
            got = obj(dpreds)
            want, _ = O.fcos_decode(preds, **kw)
            for a, b in zip(got, want):
                assert np.array_equal(a, b)
    elif name in ('RetinaFaceLoss', 'RetinaFaceDecoder'):
        gen = torch.Generator().manual_seed(84)
        cls = [torch.sigmoid(torch.randn((2, p, p, 3, 1), generator=gen) - 1.5) for p in (32, 16, 8)]
        reg = [torch.randn((2, p, p, 3, 4), generator=gen) * 0.2 for p in (32, 16, 8)]
        preds = synth.make_tie_free([cls, reg], min_score=0.3)
        ann = synth.make_annotations(2, 10, 256, 1, seed=85)
        dpreds = synth.to_device(preds, 'cuda')
        if name == 'RetinaFaceLoss':
            with torch.no_grad():
                got = obj.cuda()(dpreds, ann.cuda())
                want = O.retinaface_loss(preds, ann, **kw)
            for k in ('cls_loss', 'reg_loss'):
                np.testing.assert_allclose(got[k].item(), want[k].item(), rtol=1e-5)
        else:
            got = obj(dpreds)
            want, _ = O.retinaface_decode(preds, **kw)
            for a, b in zip(got, want):
                assert np.array_equal(a, b)
    else:   # DETRDecoder / DINODETRDecoder
        gen = torch.Generator().manual_seed(86)
        sizes = [[300, 400], [256, 512]]
        reg = torch.cat([torch.rand((2, 120, 2), generator=gen),
                         torch.rand((2, 120, 2), generator=gen) * 0.4], dim=-1)
        if name == 'DETRDecoder':
            cls = torch.randn((1, 2, 120, kw['num_classes'] + 1), generator=gen) * 2
            got = obj([cls.cuda(), reg[None].cuda()], sizes)
            want, _ = O.query_decode(cls[-1], reg, sizes, 'softmax',
                                     prob_fn=lambda x: torch.softmax(x.cuda(), dim=2).cpu(), **kw)
        else:
            cls = torch.randn((2, 120, 20), generator=gen) * 2 - 1
            got = obj({'pred_logits': cls.cuda(), 'pred_boxes': reg.cuda()}, sizes)
            want, _ = O.query_decode(cls, reg, sizes, 'sigmoid',
                                     prob_fn=lambda x: torch.sigmoid(x.cuda().float()).cpu(), **kw)
        for a, b in zip(got, want):
            assert np.array_equal(a, b)
