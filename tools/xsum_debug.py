"""Diagnostic: where does the large-domain Field.sum pass leave the serial float accumulation?  (RLFC_XS_DIAG build)"""
import re, subprocess, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import rlfluidcontrol_b200 as R
    from test_exact_sum import adversarial
    rng = np.random.default_rng(11)
    res = 128
    with R.AFCCylinderBatch(1, init_state=None, resolution=res, x_lengths=16, y_lengths=8, t_step=float(np.float32(0.18) / np.float32(res))) as env:
        n, m = env.n, env.m
        p = adversarial(1, n * m, rng).astype(np.float32)
        p = adversarial(3, n * m, rng).astype(np.float32).reshape(n, m)
        env.set_fields(0, None, None, p)
        got = env.field_sum()[0]
        np.save("/tmp/xs_p.npy", p)
        print("RESULT", got)
else:
    out = subprocess.run([sys.executable, __file__, "child"], capture_output=True, text=True).stdout
    p = np.load("/tmp/xs_p.npy")
    v = p[1:-1, 1:-1].reshape(-1)
    cs = np.cumsum(v, dtype=np.float32)          # sequential float32 accumulation
    print("true total", cs[-1], [l for l in out.splitlines() if l.startswith("RESULT")])
    bad = 0
    for l in out.splitlines():
        mm = re.match(r"xsblk (\d+) bits ([0-9a-f]+) puremask ([0-9a-f]+)", l)
        if mm:
            b0 = int(mm.group(1)); bits = int(mm.group(2), 16)
            want = np.float32(0) if b0 == 0 else cs[b0 * 1024 - 1]
            got = np.array([bits], np.uint32).view(np.float32)[0]
            if got != want and bad < 6:
                print("block", b0, "got", got, "want", want, "puremask", mm.group(3)); bad += 1
    print("blocks off:", bad)
    bad = 0
    for l in out.splitlines():
        mm = re.match(r"xsbat (\d+) bits ([0-9a-f]+)", l)
        if mm:
            b = int(mm.group(1)); bits = int(mm.group(2), 16)
            want = np.float32(0) if b == 0 else cs[b * 1024 - 1]
            got = np.array([bits], np.uint32).view(np.float32)[0]
            if got != want and bad < 4:
                print("batch", b, "got", got, "want", want); bad += 1
                if bad == 1:
                    seg = v[(b - 1) * 1024:b * 1024]
                    print("  previous batch", b - 1, "start", cs[(b - 1) * 1024 - 1], "min/max running", np.min(cs[(b-1)*1024:b*1024]), np.max(cs[(b-1)*1024:b*1024]))
    print([l for l in out.splitlines() if l.startswith("xswalk") or l.startswith("xsrun")])
    print("xsmiss lines:", [l for l in out.splitlines() if l.startswith("xs miss")][:5])
