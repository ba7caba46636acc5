"""Small decode / encode cases for compute-sanitizer (tools/gpu_sanitize.sh): short sequences so that the
instrumented persistent kernels finish in seconds.  argv: <path> <B> [beam]"""
import sys
sys.path.insert(0, '.')
import torch
import molnextr_b200.engine as E
E.MAX_LEN, E.MAX_ATOMS = 32, 10          # 32 decode steps at most (the kernels take max_len from the config)
from molnextr_b200 import synth
from tests.helpers import seeded_features

path, B = sys.argv[1], int(sys.argv[2])
beam = int(sys.argv[3]) if len(sys.argv) > 3 else 1
if path == "gemm":
    from molnextr_b200 import _cabi
    import ctypes as C
    lib = _cabi.load()
    M, N, K = 300, 256, 192
    a, w, b = torch.randn(M, K).cuda(), torch.randn(N, K).cuda(), torch.randn(N).cuda()
    out = torch.empty(M, N).cuda()
    for epi in (0, 1, 3):
        rc = lib.mnx_test_gemm_bf16(C.c_void_p(a.data_ptr()), C.c_void_p(w.data_ptr()), C.c_void_p(b.data_ptr()), C.c_void_p(out.data_ptr()),
                                    M, N, K, epi, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0, rc
    torch.cuda.synchronize()
    print("gemm ok", float(out.abs().mean()))
    sys.exit(0)
if path in ("swin", "convnext"):
    # encoder kernels: tcgen05 GEMM with the TMA store / reduce-add epilogue, pipelined window attention (cp.async), dwconv + LN fold
    ck = synth.synthetic_checkpoint(0, "sensitised", encoder="swin_base" if path == "swin" else "convnext_base")
    eng = E.Engine(ck, max_batch=B, max_height=128, max_width=160)
    x = torch.randn((B, 3, 128, 160), generator=torch.Generator().manual_seed(3)).cuda()
    f = eng.encode(x)
    torch.cuda.synchronize()
    print(path, "encode ok", tuple(f.shape), float(f.abs().mean()))
    eng.close()
    sys.exit(0)
if path == "tiled":
    import os
    os.environ["MNX_TILE_GEMM_MIN_ROWS"] = "1"      # the register-tiled GEMM of the >= 128-row graph path on a small batch
    path = "graph"
ck = {"decoder": synth.decoder_state(0, "sensitised"), "encoder": None}
eng = E.Engine(ck, max_batch=B, max_beam=max(1, beam))
if path not in ("beam", "auto"):
    eng.set_decode_path(path)
f = seeded_features(7, B, 144).cuda()
if path == "beam":
    out = eng.decode_beam(f, beam, 2)
else:
    out = eng.decode_greedy(f)
    ai, na = eng.atom_indices(out["ids"], out["lens"])
    eng.edges(ai, na)
torch.cuda.synchronize()
print(path, B, "lens", out["lens"].flatten().tolist()[:8], "steps", eng.last_decode_steps())
eng.close()
