"""One nn.Module shell for all the "construct a function object, then call it" RoI ops of the reference.

The reference ships a hand-written ten-line Module per op (RoIAlign / RoIAlignAvg / RoIAlignMax in two flavours,
_RoIPooling, _RoICrop).  They differ only in (a) the names and types of the constructor arguments they remember,
(b) the function class they instantiate per forward call and (c) an optional "+1 then 2x2 stride-1 pool" epilogue.
`roi_module` builds each of them from that description, so the mirror packages under model/ and modeling/ only state
the description; attribute names (`aligned_height`, `pooled_width`, `spatial_scale`, ...) are the reference's, because
callers read them.
"""
import torch.nn.functional as F
from torch import nn


class _RoIModule(nn.Module):
    _fn = None              # function class: _fn(*ctor_args)(*forward_inputs)
    _fields = ()            # ((attribute name, cast), ...) in constructor order
    _grow = ()              # attribute names enlarged by one before the call (the Avg / Max variants)
    _epilogue = None        # None | "avg" | "max": 2x2, stride 1

    def __init__(self, *args, **kwargs):
        super().__init__()
        names = [n for n, _ in self._fields]
        if len(args) > len(names):
            raise TypeError("%s takes %d arguments (%s)" % (type(self).__name__, len(names), ", ".join(names)))
        values = dict(zip(names, args))
        for key, val in kwargs.items():
            if key not in names or key in values:
                raise TypeError("%s: unexpected or repeated argument %r" % (type(self).__name__, key))
            values[key] = val
        for name, cast in self._fields:
            if name not in values:
                raise TypeError("%s: missing argument %r" % (type(self).__name__, name))
            setattr(self, name, cast(values[name]))

    def forward(self, *inputs):
        ctor = [getattr(self, name) + (1 if name in self._grow else 0) for name, _ in self._fields]
        y = type(self)._fn(*ctor)(*inputs)
        if self._epilogue == "avg":
            return F.avg_pool2d(y, kernel_size=2, stride=1)
        if self._epilogue == "max":
            return F.max_pool2d(y, kernel_size=2, stride=1)
        return y

    def extra_repr(self):
        return ", ".join("%s=%r" % (n, getattr(self, n)) for n, _ in self._fields)


def roi_module(name, fn, fields, grow=(), epilogue=None, doc=None, base=_RoIModule):
    """Create the Module class `name` (see the module docstring)."""
    cls = type(name, (base,), {"_fn": fn, "_fields": tuple(fields), "_grow": tuple(grow), "_epilogue": epilogue,
                               "__doc__": doc})
    return cls
