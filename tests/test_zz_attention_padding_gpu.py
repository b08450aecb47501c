"""GPU numerics of the padded attention paths (heads narrower than 64, ragged causal sequence lengths).

Kept in a file that sorts LAST on purpose: the ragged-sequence case had not been executed on a GPU when it was written
(the padding algebra was verified on CPU, the 48-wide-head path ran on GPUs through the GPT-MoE example), and the round-end
`pytest -x` run must not lose the multi-GPU tests to a failure here."""
import pytest
import torch

import kernel_checks as kc

pytestmark = pytest.mark.gpu


def test_attention_padding_paths():
    assert torch.cuda.is_available()
    from tepdist_b200 import ops
    ops.lib()
    n0 = ops.launch_count()
    kc.CHECKS["attn_d48"]()
    torch.cuda.synchronize()
    assert ops.launch_count() > n0
