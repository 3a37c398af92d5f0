"""DDP training-step host logic on CPU: world_size-2 gloo, reference op path (the B200 op has no CPU path)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

TINY = dict(decoder_layers=3, encoder_layers=2, num_queries=12, train_num_points=128)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out, batched_loss=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from weed_instance_segmentation_b200 import synth, train
        model = build()
        if batched_loss:  # the batched criterion's host logic under DDP (grid_sample stands in for the CUDA sampler)
            import sys
            sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
            from test_criterion_host import grid_sample_sampler
            from weed_instance_segmentation_b200.criterion import convert_criterion
            convert_criterion(model, sampler=grid_sample_sampler)
        tr = train.Trainer(model, "cpu")
        assert isinstance(tr.net, torch.nn.parallel.DistributedDataParallel)
        # num_masks is averaged over the group (what accelerate.reduce does at M2F:787-793)
        crit = model.criterion
        labels = [torch.zeros(3 if rank == 0 else 7, dtype=torch.int64)]
        assert float(crit.get_num_masks(labels, torch.device("cpu"))) == 5.0
        before = [p.detach().clone() for p in model.parameters()]
        batches = [synth.collate_batch(1, 64, 96, num_classes=3, max_instances=3, seed=10 * rank + i) for i in range(2)]
        tr.step(batches[0])  # first micro-batch of the window: no optimizer step, no gradient sync
        assert all(torch.equal(a, b) for a, b in zip(before, model.parameters()))
        g_local = torch.cat([p.grad.reshape(-1) for p in model.parameters() if p.grad is not None])
        gathered = [torch.zeros_like(g_local) for _ in range(world)]
        dist.all_gather(gathered, g_local)
        assert not torch.equal(gathered[0], gathered[1])  # ranks see different images, grads not yet reduced
        tr.step(batches[1])  # second micro-batch: all-reduce + AdamW step
        flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
        gathered = [torch.zeros_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        assert torch.equal(gathered[0], gathered[1])  # replicas stay identical
        assert any(not torch.equal(a, b) for a, b in zip(before, model.parameters()))
        loss = tr.mean_loss()
        assert loss == loss and loss > 0
        out[rank] = loss
    finally:
        dist.destroy_process_group()


def build():
    from weed_instance_segmentation_b200 import train
    return train.build_model("swin_tiny_test", num_labels=3, seed=0, **TINY)


@pytest.mark.parametrize("batched_loss", [False, True], ids=["stock-loss", "batched-loss"])
def test_ddp_step_world2_gloo(batched_loss):
    pytest.importorskip("transformers")
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), out, batched_loss), nprocs=world, join=True)
        assert len(out) == world


def test_single_process_accumulation_matches_reference_loop():
    """Trainer.step reproduces /root/reference/models/mask2former/train.py:191-202 on one process."""
    pytest.importorskip("transformers")
    from weed_instance_segmentation_b200 import synth, train
    batches = [synth.collate_batch(1, 64, 96, num_classes=3, max_instances=3, seed=i) for i in range(2)]

    def reference_loop(model):
        opt = torch.optim.AdamW(model.parameters(), lr=train.LEARNING_RATE)
        model.train()
        for step, batch in enumerate(batches):
            torch.manual_seed(100 + step)
            out = model(pixel_values=batch["pixel_values"], mask_labels=batch["mask_labels"],
                        class_labels=batch["class_labels"])
            (out.loss / train.GRADIENT_ACCUMULATION).backward()
            if (step + 1) % train.GRADIENT_ACCUMULATION == 0:
                opt.step()
                opt.zero_grad()

    a, b = build(), build()
    reference_loop(a)
    tr = train.Trainer(b, "cpu", ddp=False)
    for step, batch in enumerate(batches):
        torch.manual_seed(100 + step)
        tr.step(batch)
    for pa, pb in zip(a.parameters(), b.parameters()):
        assert torch.allclose(pa, pb, atol=1e-7, rtol=1e-6)


def test_modules_keep_state_dict_keys():
    """convert_pixel_decoder shares parameters: state_dict keys and tensors are unchanged."""
    pytest.importorskip("transformers")
    from weed_instance_segmentation_b200 import modules
    m = build()
    keys = list(m.state_dict().keys())
    ptrs = {k: v.data_ptr() for k, v in m.state_dict().items()}
    n = modules.convert_pixel_decoder(m)
    assert n == TINY["encoder_layers"]
    assert list(m.state_dict().keys()) == keys
    assert all(v.data_ptr() == ptrs[k] for k, v in m.state_dict().items())
    assert modules.convert_pixel_decoder(m) == 0  # idempotent
    with pytest.raises(ValueError):
        modules.MSDeformAttn(250, 8, 3, 4)


def test_batched_criterion_in_the_training_step_matches_stock_loss():
    """Trainer.step with the converted criterion (grid_sample sampler injected: the CUDA sampler has no CPU path)
    gives the stock criterion's loss and gradients for the same generator state, with uint8 masks as well."""
    import copy

    from test_criterion_host import grid_sample_sampler
    from weed_instance_segmentation_b200 import synth, train
    from weed_instance_segmentation_b200.criterion import convert_criterion

    torch.set_num_threads(4)
    stock = build()
    mine = copy.deepcopy(stock)
    convert_criterion(mine, sampler=grid_sample_sampler)
    b_f32 = synth.collate_batch(2, 64, 96, num_classes=3, max_instances=3, seed=4)
    b_u8 = synth.collate_batch(2, 64, 96, num_classes=3, max_instances=3, seed=4, mask_dtype=torch.uint8)
    assert all(torch.equal(a, b.float()) for a, b in zip(b_f32["mask_labels"], b_u8["mask_labels"]))
    results = []
    for model, batch in ((stock, b_f32), (mine, b_f32), (mine, b_u8)):
        model.zero_grad(set_to_none=True)
        tr = train.Trainer(model, "cpu", ddp=False)
        torch.manual_seed(77)
        loss = tr.step(tr.prefetch(batch))  # prefetch is the identity without a CUDA device
        grads = torch.cat([p.grad.reshape(-1) for p in model.parameters() if p.grad is not None])
        results.append((float(loss), grads.clone()))
    for loss, grads in results[1:]:
        assert abs(loss - results[0][0]) <= 1e-5 * abs(results[0][0])
        scale = results[0][1].abs().max()
        assert (grads - results[0][1]).abs().max() <= 2e-4 * scale
    assert results[1][0] == results[2][0] and torch.equal(results[1][1], results[2][1])  # mask dtype is immaterial
