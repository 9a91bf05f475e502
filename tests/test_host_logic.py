"""Host-side logic that needs no GPU: BN folding, config mapping, env sharding, the gloo stats reduce."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import load_golden
from marl_uavs_targets_tracking_b200 import (PMINetwork, default_config, episode_summary, fold_pmi, params_from_config,
                                             reduce_episode_stats, shard_envs)


def _folded_forward(f, x):
    """numpy fp32 evaluation of the folded network, the arithmetic the CUDA path performs."""
    H = f["hidden"]
    h0 = np.zeros((x.shape[0], 3 * H), np.float32)
    for b, (off, dim) in enumerate(((0, 5), (5, 4), (9, 3))):
        h0[:, b * H:(b + 1) * H] = x[:, off:off + dim] @ f["w0"][b * H:(b + 1) * H, :dim].T
    h0 = np.maximum(h0 + f["b0"], 0)
    h1 = np.maximum(h0 @ f["w1"].T + f["b1"], 0)
    return h1 @ f["w2"] + np.float32(f["b2"])


def test_bn_folding_matches_eval_forward():
    torch.manual_seed(0)
    for H in (32, 64, 128):
        net = PMINetwork(hidden_dim=H)
        for bn in (net.bn_comm, net.bn_obs, net.bn_boundary_state, net.bn1):
            bn.running_mean.normal_(0, 0.3)
            bn.running_var.uniform_(0.5, 1.5)
            bn.weight.data.normal_(1, 0.2)
            bn.bias.data.normal_(0, 0.1)
        net.eval()
        x = torch.randn(257, 12)
        with torch.no_grad():
            ref = net(x).squeeze(1).numpy()
        got = _folded_forward(fold_pmi(net), x.numpy())
        np.testing.assert_allclose(got, ref, rtol=0, atol=2e-6)
        f = fold_pmi(net)
        assert np.all(f["w0"][H:2 * H, 4:] == 0) and np.all(f["w0"][2 * H:, 3:] == 0)  # padding stays zero


def test_pmi_mirror_loads_reference_state_dict_names():
    g = load_golden("d10_pmi_s42")
    sd = {k[4:]: torch.tensor(g[k]) for k in g.files if k.startswith("pmi.")}
    net = PMINetwork(hidden_dim=128)
    net.load_state_dict(sd)  # strict: same keys and shapes as the reference's PMINetwork
    net.eval()
    assert isinstance(net.inference(np.zeros(12)), float)
    assert fold_pmi(sd)["hidden"] == 128


def test_params_from_config_converts_like_the_reference():
    p = params_from_config(default_config("MAAC-G"), 10, 10, 2000, 2000, 12, 200)
    assert (p.n_uav, p.m_targets, p.na, p.num_steps) == (10, 10, 12, 200)
    assert p.uav_h_max == np.pi / 6.0 and p.tgt_h_max == np.pi / 6.0  # src/environment.py:98,100
    assert (p.dc, p.dp, p.dt, p.uav_v_max, p.tgt_v_max) == (500, 200, 1, 20, 5)
    assert default_config("MAAC")["cooperative"] == 0  # src/main.py:75-76


def test_shard_envs_partitions_exactly():
    for total in (1, 7, 4096, 65536):
        for world in (1, 2, 3, 8):
            parts = [shard_envs(total, r, world) for r in range(world)]
            assert sum(c for c, _ in parts) == total
            assert all(parts[r][1] + parts[r][0] == (parts[r + 1][1] if r + 1 < world else total) for r in range(world))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    local = {"rewards": 1.5 + rank, "target_tracking_reward": 2.0 * (rank + 1), "boundary_punishment": -0.25,
             "duplicate_tracking_punishment": -1.0 * rank, "covered_sum": 10.0 + rank, "covered_max": 3.0 + 2 * rank,
             "env_steps": 100.0}
    red = reduce_episode_stats(local, device="cpu")
    q.put((rank, red))
    dist.destroy_process_group()


def test_stats_reduce_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = dict(q.get(timeout=120) for _ in range(2))
    [p.join(60) for p in procs]
    assert res[0] == res[1]
    r = res[0]
    assert r["rewards"] == 4.0 and r["target_tracking_reward"] == 6.0 and r["boundary_punishment"] == -0.5
    assert r["duplicate_tracking_punishment"] == -1.0 and r["covered_sum"] == 21.0 and r["covered_max"] == 5.0
    assert r["env_steps"] == 200.0
    s = episode_summary(r, n_uav=10)
    assert s["return"] == 4.0 / 2000 and s["average_covered_targets"] == 21.0 / 200 and s["max_covered_targets"] == 5.0


def test_reduce_is_identity_without_process_group():
    st = {"rewards": 1.0, "target_tracking_reward": 2.0, "boundary_punishment": 3.0, "duplicate_tracking_punishment": 4.0,
          "covered_sum": 5.0, "covered_max": 6.0, "env_steps": 7.0}
    assert reduce_episode_stats(st) == st


def _pmi_worker(rank, world, port, q):
    """Each rank trains the same PMI network on ITS OWN data (as examples/train_maac_g.py does per rank)."""
    import torch
    from marl_uavs_targets_tracking_b200 import PMINetwork, default_config
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)                      # identical initial weights on every rank
    net = PMINetwork(hidden_dim=32, b2_size=256)
    cfg = default_config("MAAC-R")
    cfg["pmi"].update(hidden_dim=32, b2_size=256, batch_size=64)
    torch.manual_seed(100 + rank)             # different replay data and different draws per rank
    data = torch.randn(40 * 10, 12)
    losses = [net.train_pmi(cfg, data, 10) for _ in range(3)]
    sd = {k: v.detach().cpu().numpy().copy() for k, v in net.state_dict().items()}  # (arrays: tensors would travel by fd)
    q.put((rank, sd, losses))
    dist.destroy_process_group()


def test_pmi_training_keeps_one_network_across_ranks_gloo():
    """several ranks train one policy against ONE reciprocal-reward network -- gradients are averaged before
    every optimizer step and BatchNorm statistics after every pass, so parameters, buffers and the reported loss are
    identical on all ranks although every rank sees its own data."""
    import torch
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_pmi_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = {r: (sd, ls) for r, sd, ls in (q.get(timeout=180) for _ in range(2))}
    [p.join(60) for p in procs]
    (sd0, l0), (sd1, l1) = res[0], res[1]
    assert l0 == l1
    for k in sd0:
        if k.endswith("num_batches_tracked"):
            continue
        assert np.array_equal(sd0[k], sd1[k]), k
    # and the network did move
    torch.manual_seed(0)
    from marl_uavs_targets_tracking_b200 import PMINetwork
    fresh = PMINetwork(hidden_dim=32, b2_size=256).state_dict()
    assert not np.array_equal(fresh["fc1.weight"].numpy(), sd0["fc1.weight"])


def test_actor_critic_checkpoints_have_the_reference_layout(tmp_path):
    """BatchedActorCritic.save / load write actor/actor_weights_N.pth and critic/critic_weights_N.pth as
    {'model_state_dict', 'optimizer_state_dict'} with the key names of the reference's FnnPolicyNet / FnnValueNet
    (src/models/actor_critic.py:85-112, :181-200), so the files load into the reference's own classes and back."""
    import torch
    from marl_uavs_targets_tracking_b200.rollout import BatchedActorCritic
    torch.manual_seed(1)
    a = BatchedActorCritic(12, 16, 12, 1e-3, 1e-3, 0.95, "cpu", fused=False)
    a.save(str(tmp_path), 7)
    pa, pc = tmp_path / "actor" / "actor_weights_7.pth", tmp_path / "critic" / "critic_weights_7.pth"
    assert pa.exists() and pc.exists()
    ca, cc = torch.load(pa, map_location="cpu"), torch.load(pc, map_location="cpu")
    assert set(ca) == set(cc) == {"model_state_dict", "optimizer_state_dict"}
    assert list(ca["model_state_dict"]) == ["fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias"]
    assert list(cc["model_state_dict"]) == ["fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias"]
    try:  # where the copy of the reference exists, load the files into ITS classes
        import baseline
        if baseline.available():
            ref = baseline.import_reference()
            agent = ref.ActorCritic(state_dim=12, hidden_dim=16, action_dim=12, actor_lr=1e-3, critic_lr=1e-3, gamma=0.95,
                                    device=torch.device("cpu"))
            agent.load(str(pa), str(pc))
            for k, v in a.actor.state_dict().items():
                assert torch.equal(agent.actor.state_dict()[k], v)
    except ImportError:
        pass
    torch.manual_seed(2)
    b = BatchedActorCritic(12, 16, 12, 1e-3, 1e-3, 0.95, "cpu", fused=False)
    assert not torch.equal(b.actor.fc1.weight, a.actor.fc1.weight)
    b.load(str(pa), str(pc))
    for (k, v), (_, w) in zip(a.actor.state_dict().items(), b.actor.state_dict().items()):
        assert torch.equal(v, w), k
    for (k, v), (_, w) in zip(a.critic.state_dict().items(), b.critic.state_dict().items()):
        assert torch.equal(v, w), k


def test_fixed_point_statistic_trick_is_linear_on_the_reward_range():
    """common.cuh: sf_fx(v) = bits(float32(v + 3)) - bits(3.0f) is the reward v in [-1, 1] as a count of 2^-22, rounded
    to nearest even -- the step kernels accumulate these integers so that the episode statistics do not depend on
    which CTA stepped which environment.  Restated in numpy: linear, exact at the ends, error <= 2^-23 per value, and
    integer sums are order-independent where float sums are not."""
    rng = np.random.RandomState(5)
    v = np.concatenate([rng.uniform(-1, 1, 200000), [-1.0, 1.0, 0.0, -0.0, 2.0 ** -30, -2.0 ** -30, 0.5, -0.5]]).astype(np.float32)
    fx = (v + np.float32(3.0)).astype(np.float32).view(np.int32).astype(np.int64) - 0x40400000
    assert fx[-8] == -(1 << 22) and fx[-7] == (1 << 22) and fx[-6] == 0 and fx[-5] == 0
    assert np.array_equal(fx, np.rint(v.astype(np.float64) * (1 << 22)).astype(np.int64))
    assert np.max(np.abs(fx / float(1 << 22) - v.astype(np.float64))) <= 2.0 ** -23
    perm = rng.permutation(v.size)
    assert fx.sum() == fx[perm].sum()
    assert abs(fx.sum() / float(1 << 22) - v.astype(np.float64).sum()) <= 1e-9 * v.size
