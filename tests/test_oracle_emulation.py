"""The operand-split emulator (oracle/emulate_operand_split.py) is what candidate tensor-core precision schemes are judged by before a
kernel exists: pin its quantisers on hand-computed values."""
import torch

from oracle import emulate_operand_split as E


def test_fp8_quantiser_rounds_and_saturates():
    x = torch.tensor([0.0, 1.0, 1.0625, 300.0, 1e6, -3.3])
    got = E.q8(x, E.E4, 0)
    # e4m3: 3 mantissa bits (1.0625 -> 1.0, ties to even), saturation at the clamp (440 -> 448 is e4m3's largest finite value)
    assert got.tolist() == [0.0, 1.0, 1.0, 288.0, 448.0, -3.25]
    # a power-of-two pre-scale is undone exactly
    assert E.q8(torch.tensor([3.0e-4]), E.E4, 12).item() == torch.tensor(3.0e-4 * 4096).to(E.E4).float().item() / 4096


def test_mxfp4_quantiser_uses_one_power_of_two_per_block_and_saturates_at_six():
    # one block of 32 channels: max 7.9 -> shared exponent floor(log2(7.9)) - 2 = 0, grid {0, .5, 1, 1.5, 2, 3, 4, 6}
    x = torch.zeros(1, 32)
    x[0, :6] = torch.tensor([7.9, 5.1, 2.4, 0.7, 0.2, -1.3])
    q = E.q4(x, 1, False)
    assert q[0, :6].tolist() == [6.0, 6.0, 2.0, 0.5, 0.0, -1.5]
    # the next block has its own exponent: max 0.11 -> 2^(-4 - 2): grid step 2^-7 at the low end
    y = torch.zeros(1, 64)
    y[0, 0] = 7.9
    y[0, 32] = 0.11
    y[0, 33] = 0.011
    q = E.q4(y, 1, False)
    assert q[0, 0].item() == 6.0 and q[0, 32].item() == 0.09375 and q[0, 33].item() == 0.0078125


def test_mxfp4_blocks_run_along_the_requested_dimension():
    w = torch.zeros(2, 64, 1, 1)
    w[0, :, 0, 0] = 3.0
    w[1, :32, 0, 0] = 0.4
    w[1, 32:, 0, 0] = 100.0
    q = E.q4(w, 1, False)
    assert torch.equal(q[0].flatten(), torch.full((64,), 3.0))
    assert torch.equal(q[1, :32].flatten(), torch.full((32,), 0.375)) and torch.equal(q[1, 32:].flatten(), torch.full((32,), 96.0))
