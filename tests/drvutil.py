"""ctypes binding of oracle/_ref/libllama_drv.so (TEST / MEASUREMENT INFRASTRUCTURE): the unmodified reference
libllama's llama_decode loop (oracle/drivers/llama_drv.cpp) inside the calling process, optionally over a ggml backend
plug-in (libggml-b200.so, or the reference's own libggml-cuda.so).  Used by bench.py's product-path legs and by the
multi-device parity tests; everything that is *measured* behind it is the plug-in, the driver only feeds tokens."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
PLUGIN = os.path.join(ROOT, "llama-box_b200", "libggml-b200.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        # the driver's own dependencies (libllama, libggml, libggml-base) sit next to it
        for dep in ("libggml-base.so", "libggml.so", "libllama.so"):
            C.CDLL(os.path.join(REF_DIR, dep), mode=C.RTLD_GLOBAL)
        L = C.CDLL(os.path.join(REF_DIR, "libllama_drv.so"), mode=C.RTLD_GLOBAL)
        vp, i32, cp = C.c_void_p, C.c_int, C.c_char_p
        L.drv_open.restype = vp; L.drv_open.argtypes = [cp, cp, i32, cp, i32, i32, i32, i32, cp, cp, i32, i32]
        L.drv_decode.restype = i32; L.drv_decode.argtypes = [vp, C.POINTER(C.c_int32), i32, i32]
        L.drv_logits.restype = C.POINTER(C.c_float); L.drv_logits.argtypes = [vp, i32]
        L.drv_embeddings.restype = C.POINTER(C.c_float); L.drv_embeddings.argtypes = [vp, i32]
        for n in ("drv_n_vocab", "drv_n_embd", "drv_n_past", "drv_n_devices"):
            getattr(L, n).restype = i32; getattr(L, n).argtypes = [vp]
        L.drv_sync.restype = None; L.drv_sync.argtypes = [vp]
        L.drv_reset.restype = None; L.drv_reset.argtypes = [vp]
        L.drv_close.restype = None; L.drv_close.argtypes = [vp]
        L.drv_handoff_stats.restype = i32
        L.drv_handoff_stats.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_double), C.POINTER(C.c_double), i32]
        _lib = L
    return _lib


def have_drv():
    return os.path.exists(os.path.join(REF_DIR, "libllama_drv.so"))


class Drv:
    """one loaded model + context"""

    def __init__(self, gguf, plugin=PLUGIN, ngl=99, ts="", ctx=4096, ubatch=512, threads=8, fa=True, ctk="f16", ctv="f16", no_repack=True, embeddings=False):
        self.L = lib()
        self.h = self.L.drv_open(gguf.encode(), (plugin or "").encode(), ngl, ts.encode(), ctx, ubatch, threads, int(fa), ctk.encode(), ctv.encode(), int(no_repack), int(embeddings))
        if not self.h:
            raise RuntimeError("drv_open failed for " + gguf)
        self.n_vocab = self.L.drv_n_vocab(self.h)
        self.n_embd = self.L.drv_n_embd(self.h)

    def decode(self, tokens, all_logits=False):
        arr = (C.c_int32 * len(tokens))(*[int(t) for t in tokens])
        rc = self.L.drv_decode(self.h, arr, len(tokens), int(all_logits))
        if rc != 0:
            raise RuntimeError(f"llama_decode returned {rc}")

    def logits(self, i=-1):
        """a VIEW of libllama's host logits buffer (valid until the next decode); synchronises the backend"""
        p = self.L.drv_logits(self.h, i)
        return np.ctypeslib.as_array(p, shape=(self.n_vocab,))

    def embeddings(self, i=-1):
        p = self.L.drv_embeddings(self.h, i)
        return np.ctypeslib.as_array(p, shape=(self.n_embd,))

    def sync(self):
        self.L.drv_sync(self.h)

    def reset(self):
        self.L.drv_reset(self.h)

    @property
    def n_past(self):
        return self.L.drv_n_past(self.h)

    @property
    def n_devices(self):
        return self.L.drv_n_devices(self.h)

    def handoff_stats(self, reset=False):
        c, b, du, hu = C.c_int64(0), C.c_int64(0), C.c_double(0), C.c_double(0)
        if not self.L.drv_handoff_stats(self.h, C.byref(c), C.byref(b), C.byref(du), C.byref(hu), int(reset)):
            return None
        return dict(copies=c.value, bytes=b.value, device_us=du.value, host_us=hu.value)

    def close(self):
        if self.h:
            self.L.drv_close(self.h); self.h = None
