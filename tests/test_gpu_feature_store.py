"""Device-resident DINO feature store: gather kernel with on-device range check, feeding the loss's
packed operand buffer directly.  Reference behaviour: src/open_clip_train/main.py:693-741, train.py:250-280."""
import pytest
import torch

from gpu_util import make_args, synth

pytestmark = pytest.mark.gpu


def _table(n=5000, dd=192, seed=3):
    g = torch.Generator().manual_seed(seed)
    return (3.0 * torch.randn(n, dd, generator=g)).float()


def test_lookup_matches_index_select(pkg):
    tab = _table()
    store = pkg.DinoFeatureStore(tab, "cuda")
    assert tuple(store.shape) == tuple(tab.shape) and len(store) == tab.shape[0]
    g = torch.Generator().manual_seed(1)
    idx = torch.randint(0, tab.shape[0], (777,), generator=g)
    want = tab.to(torch.bfloat16)[idx]
    got = store.lookup(idx)  # CPU indices are accepted like train.py's
    assert got.dtype == torch.bfloat16 and torch.equal(got.cpu(), want)
    got32 = store.lookup(idx.cuda(), dtype=torch.float32)
    assert torch.equal(got32.cpu(), want.float())
    store.check()  # nothing out of range: no error
    # fp32 table kept as is
    store32 = pkg.DinoFeatureStore(tab, "cuda", dtype=torch.float32)
    assert torch.equal(store32.lookup(idx).cpu(), want)


def test_out_of_range_is_recorded_on_device(pkg):
    tab = _table(100, 64)
    store = pkg.DinoFeatureStore(tab, "cuda")
    idx = torch.tensor([0, 5, -1, 99, 100, 7], dtype=torch.int64)
    got = store.lookup(idx).cpu()
    assert torch.equal(got[[0, 1, 3, 5]], tab.to(torch.bfloat16)[[0, 5, 99, 7]])
    assert float(got[2].abs().max()) == 0.0 and float(got[4].abs().max()) == 0.0  # bad rows are zeros, not OOB reads
    with pytest.raises(ValueError, match=r"Out-of-range indices: min=-1, max=100, feats_rows=100"):
        store.check()
    store.reset_status()
    store.lookup(torch.tensor([1, 2, 3]))
    store.check()


def test_drop_in_handle_and_loss_equivalence(pkg):
    """`args._precomputed_dino[indices].to(device, non_blocking=True)` (train.py:280) works unchanged with the store,
    and the loss fed by the lazy handle equals the loss fed by the gathered tensor bit for bit."""
    B, D, Dd = 512, 128, 192
    img, txt, _ = synth(4, B, D, Dd)
    tab = _table(4000, Dd)
    store = pkg.DinoFeatureStore(tab, "cuda")
    g = torch.Generator().manual_seed(9)
    indices = torch.randint(0, 4000, (B,), generator=g)
    rows = store[indices].to("cuda", non_blocking=True)
    assert isinstance(rows, pkg.DinoRows) and rows.size(-1) == Dd and tuple(rows.shape) == (B, Dd)
    args = make_args(use_projection=True)
    outs = []
    for dino in (rows, store.lookup(indices)):
        loss = pkg.ClipLossWithDINOEnhancements()
        torch.manual_seed(5)
        loss.init_proj(D, Dd, "cuda", "mlp")
        im = img.cuda().requires_grad_(True)
        tx = txt.cuda().requires_grad_(True)
        sc = torch.tensor(20.0, device="cuda", requires_grad=True)
        out = loss(im, tx, sc, dino, args, output_dict=True)
        out["total_loss"].backward()
        outs.append((out["total_loss"].detach().clone(), out["soft_loss"].detach().clone(), im.grad.clone(),
                     tx.grad.clone()))
    for a, b in zip(*outs):
        assert torch.equal(a, b)
    # and against the CPU gather the reference does
    ref_rows = tab[indices].to(torch.bfloat16).cuda()
    assert torch.equal(rows.materialize(), ref_rows)
