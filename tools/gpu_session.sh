cd /root/repo
timeout 600 python -m pytest tests/test_frontend.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python - <<'PY'
import torch, sys
sys.path.insert(0, "/root/repo")
from mumpy_b200 import ops
def t(fn):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 100
fr = torch.randint(0, 256, (64, 480, 854, 3), dtype=torch.uint8, device="cuda")
mid = ops.resize_u8(fr, 480, 224)
print("horizontal only %.1f us" % t(lambda: ops.resize_u8(fr, 480, 224)))
print("vertical only %.1f us" % t(lambda: ops.resize_u8(mid, 224, 224)))
print("both %.1f us" % t(lambda: ops.resize_u8(fr, 224, 224)))
sl = fr[1:3]
print("slice equal", torch.equal(ops.resize_u8(sl, 224, 224), ops.resize_u8(sl.clone(), 224, 224)))
PY
