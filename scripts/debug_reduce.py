import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ph_core_b200 as ph
from ph_core_b200 import DeviceNArray as D, sharding as S
ph.init(0); lib = ph.load()
S.comm_init(None)
stream = torch.cuda.ExternalStream(lib.ph_stream(), device=torch.device("cuda", 0))
dev = torch.device("cuda", 0)
for rows in (3, 1000):
    INNER = 1000 * 1000
    x = D([rows, 1000, 1000], np.float32)
    exact = torch.zeros((), dtype=torch.int64, device=dev)
    with torch.cuda.stream(stream):
        gen = torch.Generator(device=dev)
        for i in range(rows):
            gen.manual_seed(12345 + i)
            row = torch.randint(-8, 9, (INNER,), generator=gen, device=dev, dtype=torch.int32)
            if i == 1:
                row[777] = 99
            exact += row.sum(dtype=torch.int64)
            rowf = row.to(torch.float32)
            ph.check(lib.ph_d2d(x.ptr + i * INNER * 4, rowf.data_ptr(), INNER * 4))
    torch.cuda.synchronize()
    h = x.to_host()
    print(rows, "exact", int(exact.item()), "numpy f64 sum of device buffer", h.sum(dtype=np.float64), "argmax", int(np.argmax(h.reshape(-1))))
    print("  plain sum", x.sum(), "plain argmax", x.argmax())
    print("  sharded sum", S.reduce_full_sharded(x, "sum", 0), "sharded argmax", S.reduce_full_sharded(x, "argmax", 0),
          "sharded max", S.reduce_full_sharded(x, "max", 0))
