"""GPU: each CUDA kernel against a plain fp32 PyTorch reference of the same op, through the C ABI."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _bf16(shape, seed, scale=1.0, dev="cuda"):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(torch.bfloat16).to(dev)


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 512, 128), (200, 288, 1176), (1000, 1280, 1280),
                                   (8192, 1280, 3456), (513, 3584, 5120), (16384, 5120, 1280)])
def test_linear_matches_fp32(M, N, K):
    from video_r2_b200 import ops
    a, w = _bf16((M, K), 1), _bf16((N, K), 2, 0.05)
    bias = torch.randn(N, generator=torch.Generator().manual_seed(3)).cuda()
    out = ops.linear(a, w, bias)
    ref = a.float() @ w.float().T + bias
    err = (out - ref).abs().max().item()
    # fp32 accumulation of exact bf16 products: only summation order differs
    assert err < 2e-3 * math.sqrt(K / 64), (M, N, K, err)
    out_b = ops.linear(a, w, bias, gelu=True, out_dtype=torch.bfloat16)
    ref_b = torch.nn.functional.gelu(ref)
    assert (out_b.float() - ref_b).abs().max().item() < 0.03 * max(1.0, ref_b.abs().max().item() / 4)


def test_rmsnorm_matches_fp32():
    from video_r2_b200 import ops
    x = torch.randn(777, 1280, generator=torch.Generator().manual_seed(1)).cuda() * 3
    w = (1 + 0.1 * torch.randn(1280, generator=torch.Generator().manual_seed(2))).cuda()
    out = ops.rmsnorm(x, w, 1e-6)
    ref = w * (x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-6))
    assert (out.float() - ref).abs().max().item() <= ref.abs().max().item() * 2 ** -8


def _attn_ref(qkv, cu, heads):
    N = qkv.shape[0]
    q, k, v = qkv.float().reshape(N, 3, heads, 80).unbind(1)
    out = torch.empty(N, heads, 80, device=qkv.device)
    for i in range(len(cu) - 1):
        s, e = cu[i], cu[i + 1]
        if e == s:
            continue
        att = torch.einsum("qhd,khd->hqk", q[s:e], k[s:e]) / math.sqrt(80)
        out[s:e] = torch.einsum("hqk,khd->qhd", att.softmax(-1), v[s:e])
    return out.reshape(N, heads * 80)


@pytest.mark.parametrize("lens", [[64] * 6, [4, 16, 64, 32, 8, 48, 12], [1024], [1008, 720], [100, 1, 63, 65, 129]])
def test_attention_varlen_matches_fp32(lens):
    from video_r2_b200 import ops
    heads = 16
    cu = [0]
    for L in lens:
        cu.append(cu[-1] + L)
    qkv = _bf16((cu[-1], 3 * heads * 80), 7)
    out = ops.attention(qkv, cu, heads)
    ref = _attn_ref(qkv, cu, heads)
    err = (out.float() - ref).abs().max().item()
    assert err < 2e-2, err  # P and the output are rounded to bf16


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("lens,scale", [([64] * 6, 1.0), ([4, 16, 64, 32, 8, 48, 12], 1.0), ([1024] * 3, 1.0), ([1008, 720], 1.0),
                                        ([100, 1, 63, 65, 129], 1.0), ([2048, 300, 1024, 7], 1.0), ([196, 196, 784, 128], 1.0),
                                        ([1024, 1024], 3.0), ([640], 1.0), ([128, 129, 127], 1.0), ([64] * 300, 1.0), ([128] * 37, 2.0),
                                        ([16, 48, 64, 32, 32, 64, 128, 64, 60, 4] * 31, 1.0), ([32] * 9, 1.0)])
def test_attention_shapes(dtype, lens, scale):
    """Aligned, ragged and long segments, odd tile counts, many items per CTA, and scores whose running maximum keeps
    growing (scale 3: the lazily raised reference maximum has to rescale O in TMEM), in both 16-bit formats."""
    from video_r2_b200 import ops
    heads = 16
    cu = [0]
    for L in lens:
        cu.append(cu[-1] + L)
    g = torch.Generator().manual_seed(17)
    qkv = (torch.randn(cu[-1], 3 * heads * 80, generator=g) * scale).to(dtype).cuda()
    out = ops.attention(qkv, cu, heads)
    ref = _attn_ref(qkv, cu, heads)
    err = (out.float() - ref).abs().max().item()
    assert err < (2.5e-2 if dtype == torch.bfloat16 else 4e-3) * scale, (dtype, lens, err)


def test_attention_slice_alone_equals_slice_in_batch():
    """A slice encoded alone and inside a batch gives the same bits when its rows fall on the same 128-row tiles."""
    from video_r2_b200 import ops
    heads = 16
    g = torch.Generator().manual_seed(19)
    qkv = torch.randn(3 * 1024, 3 * heads * 80, generator=g).to(torch.float16).cuda()
    whole = ops.attention(qkv, [0, 1024, 2048, 3072], heads)
    for i in range(3):
        alone = ops.attention(qkv[i * 1024:(i + 1) * 1024].contiguous(), [0, 1024], heads)
        assert torch.equal(alone, whole[i * 1024:(i + 1) * 1024])


# ------------------------------------------------------------------ "split" precision (hi + lo bf16 operands)
@pytest.mark.parametrize("M,N,K", [(200, 288, 1176), (1000, 1280, 1280), (4096, 1280, 3456), (513, 3584, 5120)])
def test_linear_split_matches_fp32(M, N, K):
    """(A_hi + A_lo) W^T against an fp64 reference of the UNROUNDED fp32 activations: the split operand keeps
    ~16 mantissa bits, so the error is two orders below the plain bf16 GEMM on the same input."""
    from video_r2_b200 import ops
    g = torch.Generator().manual_seed(11)
    a32 = torch.randn(M, K, generator=g).cuda()
    w = _bf16((N, K), 12, scale=0.05)
    hi, lo = ops.split_bf16(a32)
    ref = (a32.double() @ w.double().T).float()
    out_split = ops.linear(hi, w, a_lo=lo)
    out_plain = ops.linear(hi, w)
    e_split = (out_split - ref).abs().max().item()
    e_plain = (out_plain - ref).abs().max().item()
    print(f"\nsplit {e_split:.2e}  plain {e_plain:.2e}")
    assert e_split < 1e-3 * math.sqrt(K / 64), e_split   # fp32 accumulation order only
    assert e_split * 20 < e_plain
    # bf16-exact activations: lo == 0 and the two kernels agree bit for bit
    z = torch.zeros_like(hi)
    assert torch.equal(ops.linear(hi, w, a_lo=z), out_plain)


@pytest.mark.parametrize("lens", [[64] * 6, [4, 16, 64, 32, 8, 48, 12], [1008, 720], [100, 1, 63, 65, 129]])
def test_attention_split_matches_fp32(lens):
    from video_r2_b200 import ops
    heads = 16
    cu = [0]
    for L in lens:
        cu.append(cu[-1] + L)
    g = torch.Generator().manual_seed(13)
    qkv32 = torch.randn(cu[-1], 3 * heads * 80, generator=g).cuda()
    hi, lo = ops.split_bf16(qkv32)
    o_hi, o_lo = ops.attention(hi, cu, heads, qkv_lo=lo)
    ref = _attn_ref(qkv32, cu, heads)
    err = (o_hi.float() + o_lo.float() - ref).abs().max().item()
    err_plain = (ops.attention(hi, cu, heads).float() - _attn_ref(qkv32, cu, heads)).abs().max().item()
    print(f"\nsplit {err:.2e}  plain {err_plain:.2e}")
    assert err < 3e-4, err


# ------------------------------------------------------------------ "f16" precision (fp16 activations x bf16 weights)
@pytest.mark.parametrize("M,N,K", [(200, 288, 1176), (1000, 1280, 1280), (4096, 1280, 3456), (513, 3584, 5120)])
def test_linear_f16_activations(M, N, K):
    """tcgen05 kind::f16 with fp16 operands (the bf16 weight converts exactly): against fp64 of the same operands the
    error is fp32 accumulation only; against the unrounded activations it is ~8x below the bf16 GEMM."""
    from video_r2_b200 import ops
    g = torch.Generator().manual_seed(11)
    a32 = torch.randn(M, K, generator=g).cuda()
    w = _bf16((N, K), 12, scale=0.05)
    a16 = a32.to(torch.float16)
    w16 = w.to(torch.float16)
    assert (w16.float() - w.float()).abs().max().item() <= 2.0 ** -25   # exact above 2^-14, subnormal grid below
    out = ops.linear(a16, w16)
    exact = (a16.double() @ w16.double().T).float()
    assert (out - exact).abs().max().item() < 1e-3 * math.sqrt(K / 64)
    ref = (a32.double() @ w.double().T).float()
    e_f16 = (out - ref).abs().max().item()
    e_bf16 = (ops.linear(a32.to(torch.bfloat16), w) - ref).abs().max().item()
    print(f"\nf16 {e_f16:.2e}  bf16 {e_bf16:.2e}")
    assert e_f16 * 4 < e_bf16
    # fp16 output: round-to-nearest, saturating at the fp16 range instead of overflowing
    o16 = ops.linear(a16, w16, out_dtype=torch.float16)
    assert o16.dtype == torch.float16 and torch.equal(o16, out.to(torch.float16))
    big = ops.linear((a16 * 0 + 60000).to(torch.float16), (w16 * 0 + 1).to(torch.float16), out_dtype=torch.float16)
    assert torch.isfinite(big).all() and big.max().item() == 65504.0


@pytest.mark.parametrize("lens", [[64] * 6, [4, 16, 64, 32, 8, 48, 12], [1008, 720], [100, 1, 63, 65, 129]])
def test_attention_f16(lens):
    from video_r2_b200 import ops
    heads = 16
    cu = [0]
    for L in lens:
        cu.append(cu[-1] + L)
    g = torch.Generator().manual_seed(13)
    qkv32 = torch.randn(cu[-1], 3 * heads * 80, generator=g).cuda()
    q16 = qkv32.to(torch.float16)
    out = ops.attention(q16, cu, heads)
    assert out.dtype == torch.float16
    ref = _attn_ref(q16.float(), cu, heads)
    err = (out.float() - ref).abs().max().item()
    err_bf16 = (ops.attention(qkv32.to(torch.bfloat16), cu, heads).float() - _attn_ref(qkv32.to(torch.bfloat16).float(), cu, heads)).abs().max().item()
    print(f"\nf16 {err:.2e}  bf16 {err_bf16:.2e}")
    assert err < 3e-3 and err * 4 < err_bf16, (err, err_bf16)


def test_forward_multi_stores_every_destination():
    """vr2_tower_forward_multi (the fused all-gather's store): the merger's last GEMM writes each row to every
    destination buffer with bulk copies; every destination must equal the single-output forward bit for bit, rows
    outside the call's token range stay untouched."""
    import video_r2_b200 as vr2
    from oracle import tower as T
    cfg_o = T.TowerConfig(depth=2, fullatt_block_indexes=(1,))
    w = T.make_weights(cfg_o, seed=71)
    tower = vr2.VisionTower(vr2.VisionConfig(depth=2, fullatt_block_indexes=(1,)), w, device="cuda:0", precision="f16")
    for grid in ([[2, 8, 12], [1, 4, 6]], [[3, 26, 46]], [[1, 2, 2]]):
        n = sum(t * h * ww for t, h, ww in grid)
        x = torch.randn(n, 1176, generator=torch.Generator().manual_seed(72)).to(torch.bfloat16).cuda()
        want = tower(x, grid)
        for n_dst in (1, 3, 8):
            lo = 5
            bufs = [torch.full((n // 4 + 9, 3584), 7.0, dtype=torch.bfloat16, device="cuda") for _ in range(n_dst)]
            tower.forward_multi(x, grid, [b[lo:].data_ptr() for b in bufs])
            torch.cuda.synchronize()
            for b in bufs:
                assert torch.equal(b[lo:lo + n // 4], want), (grid, n_dst)
                assert (b[:lo] == 7).all() and (b[lo + n // 4:] == 7).all()
    assert tower.saturation_count() == 0
    # every precision the bench cycles through at N > 1 (`modes`) reaches the same store
    grid = [[2, 8, 12], [1, 4, 6]]
    n = sum(t * h * ww for t, h, ww in grid)
    x = torch.randn(n, 1176, generator=torch.Generator().manual_seed(73)).to(torch.bfloat16).cuda()
    for prec in ("bf16", "split", "auto"):
        tower.set_precision(prec)
        want = tower(x, grid)
        bufs = [torch.zeros((n // 4, 3584), dtype=torch.bfloat16, device="cuda") for _ in range(2)]
        tower.forward_multi(x, grid, [b.data_ptr() for b in bufs])
        torch.cuda.synchronize()
        assert all(torch.equal(b, want) for b in bufs), prec
