"""Checkpoint compatibility: load an mmdet-format ``.pth``
(``{'state_dict': ..., 'meta': {...}}`` as written by mmcv's ``save_checkpoint``, possibly with
the ``module.`` prefix of (Distributed)DataParallel) into the registered detectors by parameter
name, like ``mmcv.runner.load_checkpoint(model, path, map_location='cpu')`` at
tools/hnl_test.py:746.  Parameter names are the reference's."""
import torch


def load_state_dict(model, state_dict, strict=False, logger=None):
    """mmcv-style tolerant load: reports (does not raise on) missing / unexpected keys unless strict.
    Returns (missing_keys, unexpected_keys, shape_mismatch)."""
    own = model.state_dict()
    unexpected, mismatch, ok = [], [], {}
    for name, p in state_dict.items():
        if name not in own:
            unexpected.append(name)
        elif tuple(own[name].shape) != tuple(p.shape):
            mismatch.append((name, tuple(own[name].shape), tuple(p.shape)))
        else:
            ok[name] = p
    missing = [k for k in own if k not in ok and 'num_batches_tracked' not in k]
    msg = []
    if unexpected:
        msg.append('unexpected key in source state_dict: {}'.format(', '.join(unexpected)))
    if missing:
        msg.append('missing keys in source state_dict: {}'.format(', '.join(missing)))
    for name, a, b in mismatch:
        msg.append('size mismatch for {}: model {} vs checkpoint {}'.format(name, a, b))
    if msg and strict:
        raise RuntimeError('\n'.join(msg))
    if msg and logger is not None:
        logger.warning('\n'.join(msg))
    model.load_state_dict(ok, strict=False)          # also drops the packed device weights (repacked lazily)
    return missing, unexpected, mismatch


def load_checkpoint(model, filename, map_location='cpu', strict=False, logger=None):
    checkpoint = torch.load(filename, map_location=map_location, weights_only=False)
    if isinstance(checkpoint, dict) and 'state_dict' in checkpoint:
        state_dict = checkpoint['state_dict']
    elif isinstance(checkpoint, dict):
        state_dict = checkpoint
        checkpoint = dict(state_dict=state_dict, meta={})
    else:
        raise RuntimeError('No state_dict found in checkpoint file {}'.format(filename))
    if state_dict and all(k.startswith('module.') for k in state_dict):
        state_dict = {k[7:]: v for k, v in state_dict.items()}
    load_state_dict(getattr(model, 'module', model), state_dict, strict, logger)
    checkpoint.setdefault('meta', {})
    return checkpoint
