"""GPU: a poor man's memcheck (compute-sanitizer is closed on this pool, profiles/r2_sanitizer.txt).

Every buffer the library allocates through torch.empty / torch.zeros while a full forward runs is carved out of a larger block
whose 1 KB zones before and after the payload hold a canary pattern; after the forward (all four lanes, fp16 and fp32 modes,
fused and unfused LayerNorm) every zone must be untouched -- an out-of-bounds WRITE of any kernel within 1 KB of a buffer fails
the test -- and the result must equal the unguarded run bit for bit (a kernel reading beyond a buffer would see the canaries)."""
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu

GUARD = 1024          # bytes before and after each payload
CANARY = 0x5A


class _GuardedTorch:
    """Stands in for the `torch` module inside mumpy_b200.ops and the mirror modules: empty / zeros allocate guarded blocks."""

    def __init__(self):
        self.blocks = []

    def __getattr__(self, name):
        return getattr(torch, name)

    def _alloc(self, shape, dtype, device, zero):
        if isinstance(shape, int):
            shape = (shape,)
        shape = tuple(shape)
        n = 1
        for s in shape:
            n *= int(s)
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        pad = (-nbytes) % 16
        raw = torch.full((GUARD + nbytes + pad + GUARD,), CANARY, dtype=torch.uint8, device=device)
        self.blocks.append((raw, nbytes))
        payload = raw[GUARD:GUARD + nbytes].view(dtype).view(shape)
        if zero:
            payload.zero_()
        return payload

    def empty(self, *shape, dtype=torch.float32, device=None, **kw):
        shape = shape[0] if len(shape) == 1 and not isinstance(shape[0], int) else shape
        return self._alloc(shape, dtype, device, False)

    def zeros(self, *shape, dtype=torch.float32, device=None, **kw):
        shape = shape[0] if len(shape) == 1 and not isinstance(shape[0], int) else shape
        return self._alloc(shape, dtype, device, True)

    def check(self):
        bad = 0
        for raw, nbytes in self.blocks:
            head, tail = raw[:GUARD], raw[GUARD + nbytes:]
            bad += int((head != CANARY).sum()) + int((tail != CANARY).sum())
        return bad


@pytest.mark.parametrize("mode,fused", [("fp16", False), ("fp16", True), ("fp32", False)])
def test_no_kernel_writes_outside_its_buffers(mode, fused):
    import mumpy_b200
    from mumpy_b200 import ops
    from mumpy_b200.models.decoder import decoder as dec_mod
    from mumpy_b200.models.encoder import multiTemporalViewEncoder as enc_mod
    enc, dec = mumpy_b200.Encoder().eval(), mumpy_b200.Decoder().eval()
    util.load_seeded(enc)
    util.load_seeded(dec)
    enc, dec = enc.cuda(), dec.cuda()
    x = util.seeded_input((2, 3, 3, 224, 224), 3).cuda()
    mumpy_b200.set_precision(mode)
    was, was_w = ops.FUSED_LN, ops.FUSED_LN_WIDTHS
    ops.set_fused_ln(fused, widths=(96, 128, 192, 256, 384, 512))
    guarded = _GuardedTorch()
    patched = [(m, m.torch) for m in (ops, dec_mod, enc_mod)]
    try:
        with torch.no_grad():
            ref = mumpy_b200.forward(enc, dec, x)[0].clone()           # packs the weights, unguarded
            for m, _ in patched:
                m.torch = guarded
            out = mumpy_b200.forward(enc, dec, x)[0].clone()
        torch.cuda.synchronize()
    finally:
        for m, orig in patched:
            m.torch = orig
        ops.set_fused_ln(was, widths=was_w)
        mumpy_b200.set_precision(ops.DEFAULT_PRECISION)
    assert len(guarded.blocks) > 300, "the guard hook saw only %d allocations" % len(guarded.blocks)
    assert guarded.check() == 0, "a kernel wrote outside one of the library's buffers"
    assert torch.equal(out, ref)
