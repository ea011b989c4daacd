"""CPU-only, world_size 2 over gloo: the atlas builder's distributed bookkeeping (subject sharding,
all-reduced image gradient, loss reduction) gives the same atlas and losses as one process.
The per-batch step is replaced by a small pure-torch stand-in (the real step needs a GPU); the
sharding / accumulation / collective code under test is the product's."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_builder(world_size, rank, data, batch_size=2, lr_pose=0.5):
    from lagomorph_b200.atlas import LDDMMAtlasBuilder

    class Standin(LDDMMAtlasBuilder):
        # quadratic stand-in for lddmm_step: same signature, same accumulation contract
        def lddmm_step(self, m, img, need_image_grad=True):
            m = m.detach().requires_grad_(True)
            self.I.requires_grad_(need_image_grad)
            pred = self.I + m[:, :1]
            reg = self.reg_weight * (m * m).sum() / img.numel()
            loss = ((pred - img) ** 2).sum() / img.numel() + reg
            grads = torch.autograd.grad(loss, [m, self.I] if need_image_grad else [m])
            with torch.no_grad():
                if need_image_grad:
                    self.I_grad_acc += grads[1]
                nf = img.shape[0] / self.num_subjects
                m = m.detach().add_(grads[0], alpha=-self.learning_rate_pose)
            return m, (loss * nf).detach(), (reg * nf).detach()

    return Standin(data, num_epochs=3, batch_size=batch_size, reg_weight=0.1, learning_rate_pose=lr_pose,
                   learning_rate_image=0.3, device="cpu", world_size=world_size, rank=rank,
                   metric=object())


def _worker(rank, world_size, port, data, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    b = _make_builder(world_size, rank, data)
    I, ms = b.run()
    if rank == 0:
        torch.save({"I": I, "losses": b.epoch_losses, "regs": b.epoch_reg_terms}, out)
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_ranks_match_one(tmp_path):
    torch.manual_seed(1)
    data = torch.randn(8, 1, 6, 5)
    # Two ranks x batch 2 see, per iteration, the same 4 subjects as one rank x batch 4; the image
    # gradient is averaged over ranks (lddmm.py:294-295) and the per-subject momentum gradient scales
    # with 1/batch (loss / img.numel(), lddmm.py:313), hence the doubled pose learning rate.
    # (subjects are assigned in the reference's DistributedSampler order: seed-0 permutation, strided)
    from lagomorph_b200.atlas import shard_indices
    r0, r1 = shard_indices(8, 2, 0), shard_indices(8, 2, 1)
    assert sorted(r0 + r1) == list(range(8)) and r0 != [0, 2, 4, 6]
    perm = r0[0:2] + r1[0:2] + r0[2:4] + r1[2:4]
    single = _make_builder(1, 0, data[perm], batch_size=4, lr_pose=1.0)
    I1, _ = single.run()
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), data, out), nprocs=2, join=True)
    r = torch.load(out)
    assert torch.allclose(r["I"], I1, atol=1e-6)
    assert torch.allclose(torch.tensor(r["losses"]), torch.tensor(single.epoch_losses), atol=1e-6)
    assert torch.allclose(torch.tensor(r["regs"]), torch.tensor(single.epoch_reg_terms), atol=1e-6)


# ---- affine atlas (lagomorph_b200/affine_atlas.py) ------------------------------------------------------
def _fake_affine_interp(I, A, T):
    # differentiable stand-in with the right shapes: per-subject gain and offset from (A, T)
    gain = A.diagonal(dim1=1, dim2=2).mean(1).view(-1, 1, 1, 1)
    off = 0.1 * T.sum(1).view(-1, 1, 1, 1) + 0.05 * A.sum((1, 2)).view(-1, 1, 1, 1)
    return I * gain + off


_AFF_KW = dict(num_epochs=3, batch_size=2, learning_rate_A=0.05, learning_rate_T=0.5, learning_rate_I=0.3,
               reg_weightA=0.1, reg_weightT=0.05, device="cpu", _interp=_fake_affine_interp)


def _affine_worker(rank, world_size, port, data, As, Ts, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    from lagomorph_b200.affine_atlas import affine_atlas
    I, A, T, el, il = affine_atlas(data, As.clone(), Ts.clone(), world_size=world_size, rank=rank, **_AFF_KW)
    if rank == 0:
        torch.save({"I": I, "A": A, "T": T, "el": el}, out)
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_affine_atlas_two_ranks_match_one(tmp_path):
    """2 ranks x 4 subjects (batches of 2) == one process over the same subjects in DistributedSampler
    order (seed-0 permutation strided by rank, lddmm.py:163-178): same atlas, poses and epoch losses. With
    image_update_freq = 0 the single process averages the image gradient over its 4 batches, the two
    ranks over 2 batches each and then over the ranks: the same number."""
    from lagomorph_b200.affine_atlas import affine_atlas
    torch.manual_seed(2)
    S = 8
    data = torch.randn(S, 1, 6, 5)
    As, Ts = 0.1 * torch.randn(S, 2, 2), 0.3 * torch.randn(S, 2)
    from lagomorph_b200.atlas import shard_indices
    perm = shard_indices(S, 2, 0) + shard_indices(S, 2, 1)
    I1, A1, T1, el1, _ = affine_atlas(data[perm], As[perm].clone(), Ts[perm].clone(), **_AFF_KW)
    inv = torch.argsort(torch.tensor(perm))
    out = str(tmp_path / "aff.pt")
    mp.spawn(_affine_worker, args=(2, _free_port(), data, As, Ts, out), nprocs=2, join=True)
    r = torch.load(out)
    assert torch.allclose(r["I"], I1, atol=1e-6)
    assert torch.allclose(r["A"], A1[inv], atol=1e-6) and torch.allclose(r["T"], T1[inv], atol=1e-6)
    assert torch.allclose(torch.tensor(r["el"]), torch.tensor(el1), atol=1e-6)


def test_checkpoint_roundtrip(tmp_path):
    """save() after two epochs, load() into a fresh builder, continue: same atlas and losses as
    three epochs in one go (fields of the reference's HDF5 checkpoint, lddmm.py:238-285)."""
    torch.manual_seed(3)
    data = torch.randn(6, 1, 6, 5)
    full = _make_builder(1, 0, data)
    full.num_epochs = 3
    I3, ms3 = full.run()
    a = _make_builder(1, 0, data)
    a.num_epochs = 2
    a.checkpoint_format = str(tmp_path / "ck_{epoch}.pt")
    a.run()
    b = _make_builder(1, 0, data)
    b.num_epochs = 1
    b.load(str(tmp_path / "ck_1.pt"))
    Ib, msb = b.run()
    assert torch.allclose(Ib, I3, atol=1e-7)
    assert all(torch.allclose(x, y, atol=1e-7) for x, y in zip(msb, ms3))
    assert len(b.epoch_losses) == 3 and abs(b.epoch_losses[-1] - full.epoch_losses[-1]) < 1e-7


class _FakeH5Dataset:
    def __init__(self, arr):
        self.arr, self.attrs = arr, {}
        self.shape = arr.shape

    def __setitem__(self, k, v):
        self.arr[k] = v

    def __getitem__(self, k):
        return self.arr[k]

    def __array__(self, dtype=None, copy=None):
        return self.arr if dtype is None else self.arr.astype(dtype)

    def __iter__(self):
        return iter(self.arr)


class _FakeH5File(dict):
    """just enough of h5py.File for the checkpoint code: create_dataset(data= | shape=, dtype=), item
    access, slicing, attrs, context manager; persisted as the HDF5 signature + a pickle"""
    MAGIC = b"\x89HDF\r\n\x1a\n"

    def __init__(self, name, mode):
        import pickle
        super().__init__()
        self.name, self.mode = name, mode
        if mode == "r":
            with open(name, "rb") as fh:
                assert fh.read(8) == self.MAGIC
                for k, (arr, attrs) in pickle.load(fh).items():
                    self[k] = _FakeH5Dataset(arr)
                    self[k].attrs = attrs

    def create_dataset(self, key, data=None, shape=None, dtype=None):
        import numpy as np
        self[key] = _FakeH5Dataset(np.array(data) if data is not None else np.zeros(shape, dtype))
        return self[key]

    def __enter__(self):
        return self

    def __exit__(self, *a):
        import pickle
        if self.mode == "w":
            with open(self.name, "wb") as fh:
                fh.write(self.MAGIC)
                pickle.dump({k: (d.arr, d.attrs) for k, d in self.items()}, fh)


def test_checkpoint_hdf5_layout(tmp_path, monkeypatch):
    """a name ending in .h5 goes through h5py when it is importable, in the reference's layout
    (lddmm.py:238-285: atlas, momenta + attrs['batch_sizes'], four loss lists); h5py is not in this
    image, so a dict-backed stand-in module records what the builder writes and serves it back"""
    import types
    fake = types.ModuleType("h5py")
    fake.File = _FakeH5File
    monkeypatch.setitem(sys.modules, "h5py", fake)
    torch.manual_seed(3)
    data = torch.randn(6, 1, 6, 5)
    a = _make_builder(1, 0, data)
    a.num_epochs = 2
    a.run()
    name = str(tmp_path / "atlas.h5")
    a.save(name)
    with _FakeH5File(name, "r") as h:
        assert set(h) == {"atlas", "momenta", "epoch_losses", "epoch_reg_terms", "iter_losses", "iter_reg_terms"}
        assert h["momenta"].shape[0] == 6 and list(h["momenta"].attrs["batch_sizes"]) == [m.shape[0] for m in a.ms]
        assert h["momenta"].arr.dtype.name == "float32" and len(h["epoch_losses"].arr) == 2
    b = _make_builder(1, 0, data)
    b.load(name)
    assert torch.equal(b.I0, a.I.detach().cpu())
    assert len(b.ms) == len(a.ms) and all(torch.equal(x, y.detach().cpu()) for x, y in zip(b.ms, a.ms))
    assert b.epoch_losses == [float(x) for x in a.epoch_losses] and len(b.iter_losses) == len(a.iter_losses)
    # without h5py the same name falls back to torch.save, and load() tells the formats apart
    monkeypatch.setitem(sys.modules, "h5py", None)
    a.save(name)
    c = _make_builder(1, 0, data)
    c.load(name)
    assert torch.equal(c.I0, b.I0)


def test_affine_atlas_file_roundtrip(tmp_path, monkeypatch):
    """save_affine_atlas / load_affine_atlas: the reference's fields (affine.py:579-587), HDF5 through a stand-in
    h5py module and torch.save without it"""
    import types
    from lagomorph_b200.affine_atlas import save_affine_atlas, load_affine_atlas
    torch.manual_seed(5)
    I, A, T = torch.randn(1, 1, 6, 5), torch.randn(4, 2, 2), torch.randn(4, 2)
    el, il = [3.0, 2.0], [3.5, 3.0, 2.5, 2.0]
    fake = types.ModuleType("h5py")
    fake.File = _FakeH5File
    for mod, name in ((fake, "a.h5"), (None, "b.h5"), (fake, "c.pt")):
        monkeypatch.setitem(sys.modules, "h5py", mod)
        fn = str(tmp_path / name)
        save_affine_atlas(fn, I, A, T, el, il)
        with open(fn, "rb") as fh:
            assert (fh.read(8) == _FakeH5File.MAGIC) == (mod is not None and name.endswith(".h5"))
        I2, A2, T2, el2, il2 = load_affine_atlas(fn)
        assert torch.equal(I2, I) and torch.equal(A2, A) and torch.equal(T2, T) and el2 == el and il2 == il
