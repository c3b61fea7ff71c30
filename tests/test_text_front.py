"""Text front (symbol table, sample sentences, voicepack lookup) against the live reference when it is mounted
(src/stylish_tts/lib/text_utils.py, tts/cli.py:36-81); bit-exact index work."""
import numpy as np
import pytest
import torch

import stylish_tts_b200 as st
from stylish_tts_b200 import text as T
from oracle import ref_loader


def test_token_lengths_and_padding():
    tc = T.TextCleaner(st.default_model_config().symbol)
    toks = [tc(s) for s in T.SAMPLE_PHONEMES]
    lens = [len(t) for t in toks]
    assert min(lens) == 16 and max(lens) == 71  # 16-71 tokens incl. the two `$` pads
    assert all(t[0] == 0 and t[-1] == 0 for t in toks)
    texts, lengths = T.pad_batch(toks)
    assert texts.shape == (10, 71) and lengths.tolist() == lens
    assert int(texts[0, lens[0]:].abs().sum()) == 0


@pytest.mark.skipif(not ref_loader.available(), reason="reference sources not mounted")
def test_text_cleaner_matches_reference():
    ref_loader.load()
    from stylish_tts.lib.text_utils import TextCleaner as RefCleaner

    mc = ref_loader.model_config()
    ours = T.TextCleaner(st.default_model_config().symbol)
    ref = RefCleaner(mc.symbol)
    for s in T.SAMPLE_PHONEMES + ("Hello, world! ɑːɹ ʧ ʤ", ):
        assert ours(s) == ref(s)
    lines = open(ref_loader.REF_ROOT + "/sample_dataset/training-list.txt", encoding="utf-8").read().splitlines()
    lines += open(ref_loader.REF_ROOT + "/sample_dataset/validation-list.txt", encoding="utf-8").read().splitlines()
    real = tuple(l.split("|")[1] for l in lines if len(l.split("|")) == 4)
    assert real == T.SAMPLE_PHONEMES


def test_voicepack_lookup_matches_cli_formulas():
    g = torch.Generator().manual_seed(0)
    pack = torch.randn(512, 192 + 8, generator=g)
    for n in (5, 40, 300):
        sp, pe, du = T.static_styles(pack, n)
        i = max(511, min(2, n))  # cli.py:73
        assert torch.equal(sp, pack[i:i + 1, :64]) and torch.equal(pe, pack[i:i + 1, 64:128])
        assert torch.equal(du, pack[i:i + 1, 128:192])
    idx = torch.tensor([[3, 77, 200, 9]])
    dist = torch.tensor([[0.5, 1.0, 2.0, 4.0]])
    sp, pe, du = T.dynamic_styles(pack, idx, dist)
    w = 1 / dist.numpy()
    w = (w / w.sum(axis=1))[:, :, None].astype(np.float32)  # cli.py:66-69
    p = pack.numpy()
    np.testing.assert_allclose(sp.numpy(), (p[:, :64][idx.numpy()] * w).sum(axis=1), rtol=1e-6)
    np.testing.assert_allclose(pe.numpy(), p[:, 64:128][idx.numpy()].mean(axis=1), rtol=1e-6)
    np.testing.assert_allclose(du.numpy(), p[:, 128:192][idx.numpy()].mean(axis=1), rtol=1e-6)
