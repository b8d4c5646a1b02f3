"""Development diagnostic: the GPU field squaring (bbp_test_fe op 4) against Python integers on structured inputs."""
import os, sys, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bbp_loader
pkg = bbp_loader.load()
P = 2**255 - 19
be = pkg.Backend(device=0, gens_capacity=0)
cases = []
for i in range(8):
    cases.append(1 << (32 * i))
    cases.append(0xffffffff << (32 * i))
for i in range(8):
    for j in range(i + 1, 8):
        cases.append((1 << (32 * i)) | (1 << (32 * j)))
        cases.append((0xffffffff << (32 * i)) | (0xffffffff << (32 * j)))
rnd = random.Random(3)
cases += [rnd.getrandbits(256) for _ in range(64)]
cases.append(2**256 - 1)
ab = b"".join(x.to_bytes(32, "little") for x in cases)
got4 = be.test_fe(ab, ab, 4)
got0 = be.test_fe(ab, ab, 0)
nbad = 0
for k, x in enumerate(cases):
    g4 = int.from_bytes(got4[32 * k:32 * k + 32], "little")
    g0 = int.from_bytes(got0[32 * k:32 * k + 32], "little")
    w = x * x % P
    if g4 != w or g0 != w:
        nbad += 1
        if nbad <= 12:
            print(f"case {k} x={x:#x}\n   want {w:#x}\n   sq   {g4:#x}\n   mul  {g0:#x}\n   diff(sq-want) mod p = {(g4 - w) % P:#x}")
print("bad:", nbad, "of", len(cases))
