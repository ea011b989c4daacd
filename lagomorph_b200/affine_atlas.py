"""Affine atlas building (BASELINE config 4): restatement of the driver
`affine_atlas` and of `StandardizedDataset` in lagomorph/affine.py:288-438 for in-memory data.

Kept from the reference: the per-batch step (affine_interp(I, A + 1, T) -> MSE/voxels + 0.5*reg ->
backward -> gradient steps on A and T, affine.py:355-384), the loss normalisation (:371-376,
:381,:393-394), accumulation of the atlas gradient over `image_update_freq` batches (0 = one update
per epoch) followed by an SGD step with the all-reduced gradient divided by
image_iters * world_size (:385-392,:399-405), subject sharding in DistributedSampler order.

Different (B200-first): images, A and T of a rank's shard stay resident on its GPU (the reference
moves every batch host -> device and As/Ts back, :357-359,:396-397); `.item()` is deferred to the
end of the epoch, so apart from the NCCL all_reduce of the atlas gradient an epoch has no host sync;
`add_(grad, alpha=-lr)` replaces the removed `add_(-lr, grad)` signature (:383-384).
"""
import numpy as np
import torch
import torch.distributed as dist

from .affine import affine_interp, affine_inverse
from .atlas import shard_indices


def _as_tensor_batch(dataset, ids, dtype, device):
    if torch.is_tensor(dataset):
        return dataset[ids].to(device=device, dtype=dtype)
    return torch.stack([torch.as_tensor(dataset[i]) for i in ids]).to(device=device, dtype=dtype)


def affine_atlas(dataset, As, Ts, I=None, num_epochs=1000, batch_size=50, image_update_freq=0, affine_steps=1,
                 reg_weightA=0e1, reg_weightT=0e1, learning_rate_A=1e-3, learning_rate_T=1e-2,
                 learning_rate_I=1e5, gpu=None, world_size=1, rank=0, device=None, _interp=None):
    """dataset: tensor (S, 1, X, Y[, Z]) of all subjects or an indexable of (1, X, Y[, Z]) images;
    As (S, d, d) and Ts (S, d): the affine parameters (A is stored minus the identity, as in the
    reference). Returns (I, As, Ts, epoch_losses, iter_losses) like lagomorph.affine.affine_atlas.
    `_interp`: test hook, a stand-in for affine_interp so that the sharding / accumulation / collective
    bookkeeping can be exercised on CPU over gloo (tests/test_atlas_gloo.py); never used by the product."""
    dev = torch.device(device if device is not None else ("cuda:%d" % gpu if gpu is not None else "cuda"))
    interp_fn = affine_interp if _interp is None else _interp
    if dev.type != "cuda" and _interp is None:
        raise RuntimeError("affine_atlas: the affine_interp kernels are CUDA only (no CPU fallback)")
    S = len(dataset)
    ids = shard_indices(S, world_size, rank)
    dtype = As.dtype
    imgs = _as_tensor_batch(dataset, ids, dtype, dev)               # this rank's shard, resident
    A_loc = As[ids].detach().to(dev).contiguous()
    T_loc = Ts[ids].detach().to(dev).contiguous()
    if I is None:                                                   # base image = mean of all subjects (:332-339)
        with torch.no_grad():
            I = imgs.sum(0, keepdim=True)
            cnt = torch.tensor([float(len(ids))], device=dev, dtype=dtype)
            if world_size > 1:
                dist.all_reduce(I)
                dist.all_reduce(cnt)
            I = I / cnt
    else:
        I = I.clone().to(dev)
    spatial = tuple(I.squeeze().shape)
    I = I.to(dtype).reshape((1, 1) + spatial).contiguous().requires_grad_(True)
    dim = len(spatial)
    eye = torch.eye(dim, dtype=dtype, device=dev).view(1, dim, dim)
    nvox = float(np.prod(spatial))
    L2 = lambda a, b: torch.dot(a.reshape(-1), b.reshape(-1))
    epoch_losses, iter_losses = [], []
    nloc = len(ids)

    def image_step(image_iters):
        with torch.no_grad():
            g = I.grad
            if world_size > 1:
                dist.all_reduce(g)
            I.add_(g, alpha=-learning_rate_I / (image_iters * world_size))   # SGD, no weight decay (:343)
            I.grad = None

    for epoch in range(num_epochs):
        epoch_loss = torch.zeros((), device=dev, dtype=dtype)
        it_losses = []
        image_iters = 0
        for b0 in range(0, nloc, batch_size):
            sl = slice(b0, min(b0 + batch_size, nloc))
            img = imgs[sl]
            A = A_loc[sl].clone()
            T = T_loc[sl].clone()
            n = img.shape[0]
            for affit in range(affine_steps):
                A.requires_grad_(True)
                T.requires_grad_(True)
                A.grad = None
                T.grad = None
                last = affit == affine_steps - 1         # the image gradient accumulates at the last affine step only
                Iin = I if last else I.detach()
                Idef = interp_fn(Iin, A + eye, T)
                regloss = 0.0
                if reg_weightA > 0:
                    regloss = regloss + 0.5 * reg_weightA * L2(A, A)
                if reg_weightT > 0:
                    regloss = regloss + 0.5 * reg_weightT * L2(T, T)
                loss = (torch.nn.functional.mse_loss(Idef, img, reduction="sum") * (1.0 / nvox) + regloss) / n
                loss.backward()
                with torch.no_grad():
                    li = loss.detach() * (n / S)
                    it_losses.append(li)
                    A = A.detach().add_(A.grad, alpha=-learning_rate_A)
                    T = T.detach().add_(T.grad, alpha=-learning_rate_T)
            image_iters += 1
            if image_iters == image_update_freq:
                image_step(image_iters)
                image_iters = 0
            with torch.no_grad():
                epoch_loss = epoch_loss + li
                A_loc[sl] = A
                T_loc[sl] = T
        if image_iters > 0:
            image_step(image_iters)
        if world_size > 1:
            dist.all_reduce(epoch_loss)
        epoch_losses.append(epoch_loss)
        iter_losses.extend(it_losses)
    # one host sync at the end
    epoch_losses = [float(x) for x in torch.stack(epoch_losses).cpu()] if epoch_losses else []
    iter_losses = [float(x) for x in torch.stack(iter_losses).cpu()] if iter_losses else []
    As_out, Ts_out = As.clone(), Ts.clone()
    if world_size > 1:
        # every rank returns the full parameter set: gather the shards (DistributedSampler order)
        full_A = torch.zeros((S, dim, dim), dtype=dtype, device=dev)
        full_T = torch.zeros((S, dim), dtype=dtype, device=dev)
        own = torch.zeros((S,), dtype=dtype, device=dev)
        idx = torch.as_tensor(ids, device=dev)
        full_A[idx] = A_loc
        full_T[idx] = T_loc
        own[idx] = 1.0
        # padded duplicates (S not a multiple of world_size) are owned by several ranks: average them
        dist.all_reduce(full_A)
        dist.all_reduce(full_T)
        dist.all_reduce(own)
        As_out = (full_A / own.view(S, 1, 1)).to(As.device)
        Ts_out = (full_T / own.view(S, 1)).to(Ts.device)
    else:
        As_out[ids] = A_loc.to(As.device)
        Ts_out[ids] = T_loc.to(Ts.device)
    return I.detach(), As_out, Ts_out, epoch_losses, iter_losses


class StandardizedDataset:
    """Images resampled into the atlas frame by the inverse of their affine pose (affine.py:409-438)."""

    def __init__(self, dataset, As, Ts, device="cuda"):
        self.dataset, self.As, self.Ts, self.device = dataset, As, Ts, device
        dim = Ts.shape[1]
        self.eye = torch.eye(dim, dtype=As.dtype, device=device).view(1, dim, dim)

    def __len__(self):
        return len(self.dataset)

    def __getitem__(self, idx):
        J = torch.as_tensor(self.dataset[idx]).to(self.device).unsqueeze(0)
        A = self.As[[idx], ...].to(self.device)
        T = self.Ts[[idx], ...].to(self.device)
        Ainv, Tinv = affine_inverse(A + self.eye, T)
        if J.dtype not in (torch.float32, torch.float64):
            J = J.float()
        return affine_interp(J, Ainv.to(J.dtype), Tinv.to(J.dtype)).squeeze(0)


_HDF5_MAGIC = b"\x89HDF\r\n\x1a\n"


def save_affine_atlas(filename, I, As, Ts, epoch_losses, iter_losses):
    """Write the result of affine_atlas() the way the reference's command does (affine.py:579-587): datasets
    "atlas", "A", "T", "epoch_losses", "iter_losses" in one HDF5 file -- through h5py when it is importable
    and the name ends in .h5 / .hdf5 / .hdf; otherwise the same fields through torch.save."""
    fields = {"atlas": I.detach().cpu(), "A": As.detach().cpu(), "T": Ts.detach().cpu(),
              "epoch_losses": [float(x) for x in epoch_losses], "iter_losses": [float(x) for x in iter_losses]}
    if str(filename).lower().endswith((".h5", ".hdf5", ".hdf")):
        try:
            import h5py
        except ImportError:
            h5py = None
        if h5py is not None:
            import numpy as np
            with h5py.File(filename, "w") as f:
                for k in ("atlas", "A", "T"):
                    f.create_dataset(k, data=fields[k].numpy())
                for k in ("epoch_losses", "iter_losses"):
                    f.create_dataset(k, data=np.asarray(fields[k], dtype=np.float64))
            return
    torch.save(fields, filename)


def load_affine_atlas(filename):
    """(I, As, Ts, epoch_losses, iter_losses) from a file written by save_affine_atlas() or by the
    reference (HDF5, told apart by its signature)."""
    with open(filename, "rb") as fh:
        is_hdf5 = fh.read(8) == _HDF5_MAGIC
    if is_hdf5:
        try:
            import h5py
        except ImportError as e:
            raise RuntimeError("%s is an HDF5 file and h5py is not installed" % filename) from e
        import numpy as np
        with h5py.File(filename, "r") as f:
            t = {k: torch.as_tensor(np.asarray(f[k])) for k in ("atlas", "A", "T")}
            el = [float(x) for x in np.asarray(f["epoch_losses"])]
            il = [float(x) for x in np.asarray(f["iter_losses"])]
        return t["atlas"], t["A"], t["T"], el, il
    f = torch.load(filename, map_location="cpu")
    return f["atlas"], f["A"], f["T"], list(f["epoch_losses"]), list(f["iter_losses"])
