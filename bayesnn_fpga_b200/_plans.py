"""Where execution plans (:class:`engine.Engine` objects: folded weights on the device, CUDA graphs, ctypes handles)
live: OUTSIDE the modules they were built from.

The reference's models are plain ``nn.Module`` objects that its drivers pickle whole (``torch.save(model)`` ->
``get_network(load_model=...)``, models/model_loader.py:9-17) and deep-copy; a plan stored in a module's ``__dict__``
would break both (ctypes function pointers and CUDA graphs cannot be pickled).  Plans are therefore kept in a
module-level ``WeakKeyDictionary`` keyed by the model object and are VALIDATED on every use against a fingerprint of
the model: identity of every sub-module (structure: ``blocks[i].append(MCDropout(p))`` after a forward), and the
version counter and data pointer of every parameter / buffer (in-place updates, ``load_state_dict`` on any
sub-module, fine-tuning steps, ``model.to(other_device)``).  A mismatch drops the model's plans; the next call re-plans
from the current weights.  The walk costs ~0.1 ms for ResNet-18 (250 tensors).
"""
import weakref

_CACHE = weakref.WeakKeyDictionary()          # model -> [fingerprint, {key: plan}]


def fingerprint(model):
    """Hash of (sub-module identities, parameter / buffer versions and addresses) - changes whenever a plan built from
    `model` may be stale."""
    acc = []
    push = acc.append
    stack = [model]
    while stack:
        mod = stack.pop()
        push(id(mod))
        for t in mod._parameters.values():
            if t is not None:
                push(t._version)
                push(t.data_ptr())
        for t in mod._buffers.values():
            if t is not None:
                push(t._version)
                push(t.data_ptr())
        if mod._modules:
            stack.extend(m for m in mod._modules.values() if m is not None)
    return hash(tuple(acc))


def plans_for(model):
    """The dict {key: plan} of `model`, emptied first when the model changed since the plans were built."""
    fp = fingerprint(model)
    entry = _CACHE.get(model)
    if entry is None or entry[0] != fp:
        entry = _CACHE[model] = [fp, {}]
    return entry[1]


def drop(model):
    """Forget every plan of `model` (``bnn_engine(rebuild=True)`` / explicit invalidation)."""
    _CACHE.pop(model, None)
