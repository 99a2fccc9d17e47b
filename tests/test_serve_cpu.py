"""CPU: host logic of the batched generation engine (controlar_b200/autoregressive/serve/llm.py; reference autoregressive/serve —
LLM.generate over a request queue with a CFG-aware sampler): admission order, compatibility grouping, the reference's
prompt_token_ids convention with the unconditional half, output ordering.  The GPU runner is replaced by a recording fake."""
import pytest
import torch

from controlar_b200.autoregressive.serve.llm import LLM, Request, SamplingParams, Scheduler


def _fake_runner(log):
    def run(batch, seed):
        log.append(([r.request_id for r in batch], seed))
        n = batch[0].sampling.max_tokens
        return torch.tensor([[1000 * r.request_id + t for t in range(n)] for r in batch], dtype=torch.int32)
    return run


def test_scheduler_groups_compatible_requests_fifo():
    s = Scheduler(max_images=3)
    a, b = SamplingParams(max_tokens=4), SamplingParams(max_tokens=4, top_k=100)
    for i, sp in enumerate([a, a, b, a, a, b]):
        s.add(Request(i, i, None, None, sp))
    assert [r.request_id for r in s.next_batch()] == [0, 1, 3]          # oldest first, same group, capped at 3
    assert [r.request_id for r in s.next_batch()] == [2, 5]
    assert [r.request_id for r in s.next_batch()] == [4]
    assert s.next_batch() == [] and not s.has_unfinished()


def test_generate_reference_convention_with_cfg():
    log = []
    llm = LLM(cfg_scale=4.0, runner=_fake_runner(log), max_images_per_batch=8, seed=5)
    llm.num_classes = 1000
    labels = list(range(11))
    ids = [[c] for c in labels] + [[1000] for _ in labels]
    outs = llm.generate(prompt_token_ids=ids, sampling_params=SamplingParams(max_tokens=6, top_k=2000), use_tqdm=False)
    assert [b for b, _ in log] == [list(range(8)), [8, 9, 10]] and [s for _, s in log] == [5, 6]     # 11 images -> batches of 8 + 3
    assert len(outs) == 22
    for i, o in enumerate(outs[:11]):
        assert o.outputs[0].token_ids == [1000 * i + t for t in range(6)] and o.prompt_token_ids == [i]
    for i, o in enumerate(outs[11:]):
        assert o.outputs[0].token_ids == outs[i].outputs[0].token_ids and o.prompt_token_ids == [1000]   # uncond rows mirror their partner
    with pytest.raises(ValueError):
        llm.generate(prompt_token_ids=[[1], [2]], sampling_params=SamplingParams(max_tokens=2))      # second half must be the null class


def test_queue_interface_and_mixed_groups():
    log = []
    llm = LLM(cfg_scale=1.0, runner=_fake_runner(log), max_images_per_batch=4)
    sp1, sp2 = SamplingParams(max_tokens=3), SamplingParams(max_tokens=5)
    ids = [llm.add_request(i, sp1 if i % 2 == 0 else sp2) for i in range(6)]
    got = {}
    while llm.has_unfinished_requests():
        for o in llm.step():
            got[o.request_id] = o
    assert sorted(got) == ids
    assert [b for b, _ in log] == [[0, 2, 4], [1, 3, 5]]
    assert len(got[0].outputs[0].token_ids) == 3 and len(got[1].outputs[0].token_ids) == 5
