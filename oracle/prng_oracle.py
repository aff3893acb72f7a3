"""CPU restatement of utils.PRNG and ring.CRPGenerator of the reference -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the
product (lattigo-fhe-by-go_b200/) never does.

The hash is a third-party dependency of the reference that is absent from /root/reference:
golang.org/x/crypto/blake2b, pinned at v0.0.0-20190701094942-4def268fd1a4 (go.mod:5), call site
utils/prng.go:5,25 (`blake2b.New512(key)`).  Its published algorithm is BLAKE2b (RFC 7693) with a
64-byte digest and an optional key; here it is Python's hashlib.blake2b (the BLAKE2 reference C
code), an independent implementation, so parity of the library's own BLAKE2b is pinned against it
and against the RFC 7693 appendix A vector (tests/test_prng.py).  The chain and the sampler follow
    utils/prng.go:22-72   NewPRNG / Seed / Clock / SetClock
    ring/prng.go:21-103   NewCRPGenerator / Clock
"""
import hashlib

import numpy as np


class PRNG:
    def __init__(self, key=None):  # utils/prng.go:22-28
        key = bytes(key or b"")
        if len(key) > 64:
            raise ValueError("blake2b: invalid key size")
        self.key = key
        self.hash = hashlib.blake2b(key=key, digest_size=64)
        self.clock = 0
        self.seed = b""

    def Seed(self, seed):  # :38-43  Reset keeps the key
        self.hash = hashlib.blake2b(key=self.key, digest_size=64)
        self.seed = bytes(seed)
        self.hash.update(self.seed)
        self.clock = 0

    def Clock(self):  # :51-56  Sum(nil) does not disturb the running state; Write absorbs the digest
        tmp = self.hash.digest()
        self.hash.update(tmp)
        self.clock += 1
        return tmp

    def SetClock(self, n):  # :61-72
        if self.clock > n:
            raise ValueError("error : cannot set prng clock to a previous state")
        while self.clock != n:
            self.Clock()


class CRPGenerator:
    def __init__(self, key, N, moduli):  # ring/prng.go:21-37
        self.prng = PRNG(key)
        self.N = N
        self.moduli = [int(q) for q in moduli]
        self.masks = [(1 << q.bit_length()) - 1 for q in self.moduli]

    def Seed(self, seed):
        self.prng.Seed(seed)

    def SetClock(self, n):
        self.prng.SetClock(n)

    def GetClock(self):
        return self.prng.clock

    def Clock(self):  # ring/prng.go:71-103 -> [nlimbs][N] uint64
        out = np.zeros((len(self.moduli), self.N), dtype=np.uint64)
        rb = self.prng.Clock()
        for i in range(self.N):
            for j, qi in enumerate(self.moduli):
                while True:
                    if len(rb) < 8:
                        rb = self.prng.Clock()
                    coeff = int.from_bytes(rb[:8], "big") & self.masks[j]
                    rb = rb[8:]
                    if coeff < qi:
                        break
                out[j, i] = coeff
        return out
