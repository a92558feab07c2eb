"""Byte-level mutations of valid synthetic samples, gzip-wrapped, for scripts/fuzz_shards_driver.cpp.

    python scripts/fuzz_shards_gen.py SEED OUT_DIR NUM_FILES
"""
import sys, gzip, random, os, msgpack, copy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "neurips21-self-supervised-bug-detection-and-repair_b200"))
from buglab_b200.synthetic import SyntheticBugLabGenerator
gen = SyntheticBugLabGenerator(seed=3, mean_nodes=80, min_nodes=30)
base = [msgpack.packb(gen.sample(), use_bin_type=True) for _ in range(4)]
rng = random.Random(int(sys.argv[1])); out = sys.argv[2]; os.makedirs(out, exist_ok=True)
interesting = [0x00,0x7f,0x80,0x8f,0x90,0x9f,0xa0,0xbf,0xc0,0xc1,0xc2,0xc4,0xc7,0xca,0xcb,0xcc,0xcd,0xce,0xcf,0xd0,0xd3,0xd4,0xd9,0xda,0xdb,0xdc,0xdd,0xde,0xdf,0xe0,0xff]
for f in range(int(sys.argv[3])):
    payload = b""
    for b in rng.sample(base, 2):
        b = bytearray(b)
        for _ in range(rng.choice([1,1,1,2])):
            pos = rng.randrange(len(b)); mode = rng.random()
            if mode < 0.4: b[pos] = rng.choice(interesting)
            elif mode < 0.6: b[pos] = rng.randrange(256)
            elif mode < 0.75: del b[pos: pos + rng.choice([1,2,8,64])]
            elif mode < 0.9: b[pos:pos] = bytes(rng.choice(interesting) for _ in range(rng.choice([1,2,5])))
            else: b[pos:pos+4] = b"\xff\xff\xff\xff"
        payload += bytes(b)
    data = gzip.compress(payload, 1)
    if rng.random() < 0.1: data = data[: rng.randrange(len(data))]
    open(os.path.join(out, f"f{f:05d}.gz"), "wb").write(data)
