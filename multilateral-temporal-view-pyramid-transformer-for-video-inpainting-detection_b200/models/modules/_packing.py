"""Weight packing helpers shared by the mirror modules.

Parameters stay ordinary fp32 nn.Parameters under the reference's state_dict keys; the kernel-side layouts
(bf16 copies, gathered relative-position bias, fused k/v weights, NHWC conv filters, ...) are derived lazily
and cached per module, keyed on the parameters' storage, version counter and the precision mode, so a
load_state_dict / .cuda() / in-place update invalidates them automatically.
"""
import torch

from ... import ops


def invalidate_packed(module):
    """Drops every cached kernel-side copy below `module`.  Needed after in-place updates through `.data` (p.data.copy_(w),
    EMA updates ...), which do not bump Parameter._version; load_state_dict and checkpoint.restore call it themselves.
    CUDA graphs captured earlier hold raw pointers to the old copies and must be re-captured (also after set_precision)."""
    for m in module.modules():
        m.__dict__.pop("_pack_cache", None)


class PackedModule(torch.nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.register_load_state_dict_post_hook(lambda mod, incompatible: invalidate_packed(mod))

    def _packed(self, name, params, fn):
        key = tuple((p.data_ptr(), p._version, p.device) for p in params) + (ops.precision(),)
        cache = self.__dict__.setdefault("_pack_cache", {})
        hit = cache.get(name)
        if hit is None or hit[0] != key:
            with torch.no_grad():
                hit = (key, fn())
            cache[name] = hit
            # Packing is enqueued on whatever stream is current (often one of the branch lanes of streams.py) but the packed
            # tensor is cached and may be consumed from any other stream later: finish it before anyone can enqueue a
            # consumer.  Once per weight load; never under graph capture (bench.py packs in an eager warm-up step).
            dev = next((p.device for p in params if p.is_cuda), None)
            if dev is not None and not torch.cuda.is_current_stream_capturing():
                torch.cuda.current_stream(dev).synchronize()
        return hit[1]

    def _gemm_weight(self, name, param, reshape=None):
        """(N,K) GEMM operand of `param` in the current precision (fp32 as is, or a cached bf16 copy)."""
        def make():
            w = param.detach()
            if reshape is not None:
                w = w.reshape(reshape)
            w = w.contiguous()
            return ops.cast16(w) if ops.tensor_cores() else w
        return self._packed("w:" + name, [param], make)


def as_operand(x):
    """fp32 activation -> GEMM A operand of the current precision."""
    if ops.tensor_cores() and x.dtype == torch.float32:
        return ops.cast16(x)
    return x


def require_inference(module):
    if module.training:
        raise RuntimeError("%s: libmumpy_b200 implements the inference forward only; call .eval()" % type(module).__name__)
