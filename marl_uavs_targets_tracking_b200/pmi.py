"""PMI network: host-side mirror of the reference interface and BatchNorm folding.

`PMINetwork` has the reference's constructor, parameter names (`state_dict()` keys) and
`forward` / `inference` semantics (src/models/PMINet.py:20-72), so checkpoints written by the
reference's `PMINetwork.save` load here and vice versa.  `train_pmi` (src/models/PMINet.py:74-100,
SURVEY.md section 8f row f-3) keeps the reference's sampling rule, loss and optimizer but gathers the
3 000 state pairs with one indexed load on whatever device the module lives on instead of a python
loop; the environment picks the new weights up through `pmi_version`.

`fold_pmi` turns any module / state dict with those names into the BN-folded fp32 arrays the
CUDA path consumes (include/uavsim.h: UavSimPmiWeights).
"""
import os

import numpy as np
import torch
import torch.nn as nn

BN_EPS = 1e-5  # torch.nn.BatchNorm1d default, used by the reference unchanged
_BRANCHES = (("fc_comm", "bn_comm", 5), ("fc_obs", "bn_obs", 4), ("fc_boundary_state", "bn_boundary_state", 3))


class PMINetwork(nn.Module):
    """12-d -> scalar MLP: three branch Linear+BN+ReLU over the communication (5), observation (4)
    and boundary/state (3) slices, concatenated, Linear+BN+ReLU, Linear (src/models/PMINet.py:29-62)."""

    def __init__(self, comm_dim=5, obs_dim=4, boundary_state_dim=3, hidden_dim=64, b2_size=3000):
        super().__init__()
        self.comm_dim, self.obs_dim, self.boundary_state_dim = comm_dim, obs_dim, boundary_state_dim
        self.hidden_dim, self.b2_size = hidden_dim, b2_size
        self.fc_comm = nn.Linear(comm_dim, hidden_dim)
        self.bn_comm = nn.BatchNorm1d(hidden_dim)
        self.fc_obs = nn.Linear(obs_dim, hidden_dim)
        self.bn_obs = nn.BatchNorm1d(hidden_dim)
        self.fc_boundary_state = nn.Linear(boundary_state_dim, hidden_dim)
        self.bn_boundary_state = nn.BatchNorm1d(hidden_dim)
        self.fc1 = nn.Linear(hidden_dim * 3, hidden_dim)
        self.bn1 = nn.BatchNorm1d(hidden_dim)
        self.fc2 = nn.Linear(hidden_dim, 1)
        self.optimizer = torch.optim.Adam(self.parameters(), lr=0.001)  # src/models/PMINet.py:39

    def forward(self, x):
        if isinstance(x, np.ndarray):
            x = torch.tensor(x, dtype=torch.float32)
        x = x.float()
        a, b = self.comm_dim, self.comm_dim + self.obs_dim
        parts = (torch.relu(self.bn_comm(self.fc_comm(x[:, :a]))),
                 torch.relu(self.bn_obs(self.fc_obs(x[:, a:b]))),
                 torch.relu(self.bn_boundary_state(self.fc_boundary_state(x[:, b:b + self.boundary_state_dim]))))
        return self.fc2(torch.relu(self.bn1(self.fc1(torch.cat(parts, dim=1)))))

    def inference(self, single_data):
        self.eval()
        if isinstance(single_data, np.ndarray):
            single_data = torch.tensor(single_data, dtype=torch.float32)
        if single_data.ndim == 1:
            single_data = single_data.unsqueeze(0)
        with torch.no_grad():
            return self.forward(single_data).item()

    @staticmethod
    def pair_loss(output1, output2):
        """CustomLoss (src/models/PMINet.py:10-17): mean(log(1 + e^-o1) + log(1 + e^o2))."""
        return torch.mean(torch.log(1 + torch.exp(-output1)) + torch.log(1 + torch.exp(output2)))

    def train_pmi(self, config, train_data, n_uav):
        """One PMI update (src/models/PMINet.py:74-100).  train_data [timesteps*n_uav, 12]: `b2_size` draws of
        (timestep, uav A, uav B) -- same `torch.randint` calls on the CPU generator as the reference, so a seeded run
        selects the same rows -- then minibatches of config['pmi']['batch_size'] through Adam.  Returns the mean
        |loss| like the reference.  Runs on the module's device; train_data may live anywhere."""
        self.train()
        dev = self.fc2.weight.device
        timesteps = train_data.size(0) // n_uav
        # (the reference's view() needs a multiple of n_uav rows; a ragged tail is dropped here instead of raising)
        data = train_data[:timesteps * n_uav].view(timesteps, n_uav, 12).to(dev, torch.float32)
        timestep_indices = torch.randint(low=0, high=timesteps, size=(self.b2_size,))
        uav_indices = torch.randint(low=0, high=n_uav, size=(self.b2_size, 2))
        selected = data[timestep_indices.to(dev)[:, None], uav_indices.to(dev)]  # [b2, 2, 12]
        bs = config["pmi"]["batch_size"]
        nb = self.b2_size // bs
        total = torch.zeros((), device=dev, dtype=torch.float64)  # the reference sums python floats
        # Several ranks training one policy must also share ONE reward network (each rank folds its PMI into its own
        # environment): gradients are averaged over the ranks before every optimizer step, and the BatchNorm running
        # statistics -- updated from each rank's own batches -- are averaged after the pass, so parameters, buffers
        # and Adam state stay identical on all ranks (tests/test_host_logic.py, world size 2 over gloo).
        import torch.distributed as dist
        world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        params = [p for p in self.parameters() if p.requires_grad]
        for i in range(nb):
            self.optimizer.zero_grad()
            batch = selected[i * bs:(i + 1) * bs]
            loss = self.pair_loss(self.forward(batch[:, 0]), self.forward(batch[:, 1]))
            total = total + loss.detach().abs().double()
            loss.backward()
            if world > 1:
                flat = torch.cat([p.grad.reshape(-1) for p in params])
                dist.all_reduce(flat)
                flat /= world
                o = 0
                for p in params:
                    p.grad.copy_(flat[o:o + p.numel()].view_as(p.grad))
                    o += p.numel()
            self.optimizer.step()
        if world > 1:
            with torch.no_grad():
                bufs = [b for name, b in self.named_buffers() if b.dtype.is_floating_point]
                flat = torch.cat([b.reshape(-1) for b in bufs])
                dist.all_reduce(flat)
                flat /= world
                o = 0
                for b in bufs:
                    b.copy_(flat[o:o + b.numel()].view_as(b))
                    o += b.numel()
                dist.all_reduce(total)
                total /= world
        return float(total / nb)  # one device->host sync per call instead of one per minibatch

    def save(self, save_dir, epoch_i):
        """Same file layout as the reference (src/models/PMINet.py:102-106)."""
        torch.save({"model_state_dict": self.state_dict(), "optimizer_state_dict": self.optimizer.state_dict()},
                   os.path.join(save_dir, "pmi", "pmi_weights_" + str(epoch_i) + ".pth"))

    def load(self, path):
        if path and os.path.exists(path):
            checkpoint = torch.load(path, map_location=self.fc2.weight.device)
            self.load_state_dict(checkpoint["model_state_dict"])
            self.optimizer.load_state_dict(checkpoint["optimizer_state_dict"])


def _as_state(pmi):
    sd = pmi.state_dict() if hasattr(pmi, "state_dict") else pmi
    out = {}
    for k, v in sd.items():
        out[k] = v.detach().cpu().double().numpy() if torch.is_tensor(v) else np.asarray(v, dtype=np.float64)
    return out


def fold_pmi(pmi):
    """Fold eval-mode BatchNorm (running statistics) into the preceding Linear layers.

    y = (W x + b - mean) / sqrt(var + eps) * gamma + beta  ==  (s*W) x + ((b - mean)*s + beta),
    s = gamma / sqrt(var + eps).  Folding is done in float64, the result rounded to float32.
    Returns dict(hidden, w0 [3H,5], b0 [3H], w1 [H,3H], b1 [H], w2 [H], b2 float).
    """
    sd = _as_state(pmi)
    H = sd["fc1.weight"].shape[0]
    w0 = np.zeros((3 * H, 5), np.float64)
    b0 = np.zeros(3 * H, np.float64)
    for b, (fc, bn, dim) in enumerate(_BRANCHES):
        s = sd[bn + ".weight"] / np.sqrt(sd[bn + ".running_var"] + BN_EPS)
        w0[b * H:(b + 1) * H, :dim] = sd[fc + ".weight"] * s[:, None]
        b0[b * H:(b + 1) * H] = (sd[fc + ".bias"] - sd[bn + ".running_mean"]) * s + sd[bn + ".bias"]
    s1 = sd["bn1.weight"] / np.sqrt(sd["bn1.running_var"] + BN_EPS)
    w1 = sd["fc1.weight"] * s1[:, None]
    b1 = (sd["fc1.bias"] - sd["bn1.running_mean"]) * s1 + sd["bn1.bias"]
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)  # noqa: E731
    return {"hidden": int(H), "w0": f32(w0), "b0": f32(b0), "w1": f32(w1), "b1": f32(b1),
            "w2": f32(sd["fc2.weight"].reshape(-1)), "b2": float(np.float32(sd["fc2.bias"].reshape(-1)[0]))}


def pmi_version(pmi):
    """Cheap change detector: sum of the in-place version counters of all parameters and buffers
    (optimizer steps, load_state_dict and BN running-stat updates all bump them)."""
    if not hasattr(pmi, "state_dict"):
        return id(pmi)
    return (id(pmi), sum(int(v._version) for v in pmi.state_dict(keep_vars=True).values()))
