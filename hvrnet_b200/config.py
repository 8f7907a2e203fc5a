"""A small stand-in for ``mmcv.Config.fromfile`` (mmcv is not in this image): executes a
reference-style config module (plain Python with module-level logic,
configs/faster_rcnn_r101_hrnmp_c5.py:8-32) and wraps its dicts for attribute access with the
dict methods the hot path uses (``cfg.nms_pre``, ``cfg.get``, ``hasattr``, ``.copy()``,
``.pop('type')`` - rpn_head.py:77-101, bbox_nms.py:32-33)."""
import os


class ConfigDict(dict):

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError("'ConfigDict' object has no attribute '%s'" % name)

    def __setattr__(self, name, value):
        self[name] = _wrap(value)

    def copy(self):
        return ConfigDict(dict.copy(self))


def _wrap(v):
    if isinstance(v, ConfigDict):
        return v
    if isinstance(v, dict):
        return ConfigDict({k: _wrap(x) for k, x in v.items()})
    if isinstance(v, list):
        return [_wrap(x) for x in v]
    if isinstance(v, tuple):
        return tuple(_wrap(x) for x in v)
    return v


class Config(object):

    def __init__(self, cfg_dict=None, filename=None, text=''):
        object.__setattr__(self, '_cfg_dict', _wrap(cfg_dict or {}))
        object.__setattr__(self, '_filename', filename)
        object.__setattr__(self, '_text', text)

    @staticmethod
    def fromfile(filename):
        filename = os.path.abspath(os.path.expanduser(filename))
        if not os.path.isfile(filename):
            raise FileNotFoundError('file "{}" does not exist'.format(filename))
        with open(filename, 'r') as f:
            text = f.read()
        ns = {'__file__': filename, '__name__': '__hvr_config__'}
        exec(compile(text, filename, 'exec'), ns)
        cfg = {k: v for k, v in ns.items() if not k.startswith('__') and not callable(v) and
               not isinstance(v, type(os))}
        return Config(cfg, filename, text)

    @property
    def filename(self):
        return self._filename

    @property
    def text(self):
        return self._text

    def __getattr__(self, name):
        return getattr(self._cfg_dict, name)

    def __getitem__(self, name):
        return self._cfg_dict[name]

    def __setattr__(self, name, value):
        self._cfg_dict[name] = _wrap(value)

    def __contains__(self, name):
        return name in self._cfg_dict

    def get(self, name, default=None):
        return self._cfg_dict.get(name, default)

    def __repr__(self):
        return 'Config (path: {}): {}'.format(self._filename, dict.__repr__(self._cfg_dict))
