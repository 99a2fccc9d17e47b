"""Drop-in shim: `from autoregressive.serve.llm import LLM` (reference serve/llm.py) resolves to the engine built on the persistent
decode kernel (no vLLM).  `SamplingParams` of this module replaces `vllm.SamplingParams` for the fields the serve scripts set."""
from controlar_b200.autoregressive.serve.llm import *  # noqa: F401,F403
from controlar_b200.autoregressive.serve.llm import LLM, SamplingParams, RequestOutput, CompletionOutput, Scheduler, Request  # noqa: F401
