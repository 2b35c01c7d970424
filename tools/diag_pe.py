"""PE prologue of the layer GEMM against the same product on encoded planes (diagnostic)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from nefii_b200 import ops, mlp
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    for fmt in (0, 1):
        for rows in (1, 40, 127, 300, 5000):
            x = (torch.rand(rows, 3, device=dev) * 2 - 1)
            w = torch.randn(512, 39, device=dev) / 6
            enc = mlp.encode_segments([(x, 6)])
            a = ops.split_to_planes(enc, cols_pad=64, fmt=fmt)
            b = ops.split_to_planes(w, rows_pad=512, cols_pad=64, fmt=fmt)
            o1 = torch.zeros(rows, 512, device=dev)
            o2 = torch.zeros(rows, 512, device=dev)
            ops.gemm_split_bf16(a, b, 64, 512, dst_f32=o1, f32_begin=0, f32_end=512)
            pdt = torch.float16 if fmt else torch.bfloat16
            side = (torch.zeros(rows, 512, device=dev, dtype=pdt), torch.zeros(rows, 512, device=dev, dtype=pdt))
            ops.gemm_split_bf16(None, b, 64, 512, dst_f32=o2, f32_begin=0, f32_end=512, pe_x=x, pe_n_freqs=6, pe_side=side, pe_side_col0=473,
                                pe_side_scale=0.5)
            torch.cuda.synchronize()
            ref = enc.double() @ w.double().t()
            d = (o1 - o2).abs()
            bad_rows = torch.nonzero(d.amax(1) > 1e-4).squeeze(1)
            sref = enc * 0.5
            sgot = (side[0].float() + side[1].float())[:, 473:512]
            print("fmt %d rows %5d: planes vs f64 %.2e | PE prologue vs planes max %.2e (bad rows: %s) | side err %.2e, outside side cols %.1f" % (
                fmt, rows, (o1.double() - ref).abs().max().item(), d.max().item(), bad_rows[:12].tolist(), (sgot - sref).abs().max().item(),
                (side[0].float().abs() + side[1].float().abs())[:, :473].max().item()))


if __name__ == "__main__":
    main()
