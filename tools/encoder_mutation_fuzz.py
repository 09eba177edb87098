"""Mutation fuzz of the native batch encoder's protobuf reader (cerbos_b200/csrc/cb_encode.h) under AddressSanitizer +
UndefinedBehaviorSanitizer: valid serialized CheckInput messages of the C5 / C3 workloads with bits flipped, bytes inserted,
tails cut off, or replaced by noise; every batch must be encoded or refused -- never read out of bounds.

    python tools/encoder_mutation_fuzz.py [seed] [batches per workload]

Builds tests/hostsim/hostsim.cpp (which includes the encoder) with -fsanitize=address,undefined into /tmp and re-executes
itself with libasan preloaded.  No device needed."""
import ctypes
import os
import random
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = "/tmp/cerbos_b200_asan/libhostsim_asan.so"


def main():
    if os.environ.get("CB_ASAN_CHILD") != "1":
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-fsanitize=address,undefined", "-fno-omit-frame-pointer",
                        f"-I{ROOT}/include", f"-I{ROOT}/cerbos_b200/csrc", "-o", SO, f"{ROOT}/tests/hostsim/hostsim.cpp"], check=True)
        asan = subprocess.run(["g++", "-print-file-name=libasan.so"], capture_output=True, text=True, check=True).stdout.strip()
        env = dict(os.environ, CB_ASAN_CHILD="1", LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0")
        sys.exit(subprocess.run([sys.executable, __file__] + sys.argv[1:], env=env).returncode)
    sys.path.insert(0, ROOT)
    import workloads as W
    from cerbos_b200 import wire
    lib = ctypes.CDLL(SO)
    lib.hostsim_encode.restype = ctypes.c_void_p
    r = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
    per = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    ok = refused = 0
    for wname in ("C5", "C3"):
        w = W.WORKLOADS[wname]()
        _, ft, _ = W.build(w)
        msgs = [wire.check_input(i) for i in w.inputs(w.fields(64), range(64))]
        blob = ctypes.create_string_buffer(ft.blob, len(ft.blob))
        for _ in range(per):
            batch = []
            for m in r.sample(msgs, 8):
                b = bytearray(m)
                k = r.random()
                if k < 0.3:
                    for _ in range(r.randrange(1, 4)):
                        b[r.randrange(len(b))] ^= 1 << r.randrange(8)
                elif k < 0.5:
                    b = b[: r.randrange(0, len(b))]
                elif k < 0.6:
                    i = r.randrange(len(b))
                    b[i:i] = bytes(r.randrange(256) for _ in range(r.randrange(1, 9)))
                elif k < 0.7:
                    b[r.randrange(len(b))] = 0xFF
                elif k < 0.75:
                    b = bytearray(r.randrange(256) for _ in range(r.randrange(0, 40)))
                batch.append(bytes(b))
            bufs = [(ctypes.c_char * max(len(m), 1)).from_buffer_copy(m or b"\0") for m in batch]   # exact-size heap blocks
            ptrs = (ctypes.c_void_p * len(batch))(*[ctypes.addressof(x) for x in bufs])
            lens = (ctypes.c_uint64 * len(batch))(*[len(m) for m in batch])
            dims = (ctypes.c_uint32 * 4)()
            h = lib.hostsim_encode(blob, ctypes.c_uint64(len(ft.blob)), b"default", b"", ctypes.c_int(0), ptrs, lens,
                                   ctypes.c_uint64(len(batch)), dims, ctypes.c_uint32(r.choice([1, 3])))
            if h:
                ok += 1
                lib.hostsim_encoded_free(ctypes.c_void_p(h))
            else:
                refused += 1
    print(f"batches encoded {ok}, refused {refused}, no sanitizer report")


if __name__ == "__main__":
    main()
