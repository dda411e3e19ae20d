"""TEST INFRASTRUCTURE ONLY: loop restatement of the future-window construction of
cadm/samplers/model_sample_processor.py:60-92, to check cadm_b200.samplers.ModelSampleProcessor at sizes and shapes the
recorded reference scenarios (tests/golden/recorded/sampler_golden.npz, produced by the unmodified reference) do not cover.
Pinned by those scenarios in tests/test_samplers.py."""
import numpy as np


def future_windows_loops(observations, actions, F):
    """One (already long enough or not) path -> concat_obs, concat_act, concat_next_obs, concat_bool, element by element.
    Row t, step i holds observation t+i, action t+i and observation t+i+1, zero where the path has ended; a step is valid
    when transition t+i really happened; the reference clears the whole first row as well (`concat_bool[-0]`)."""
    L0 = observations.shape[0]
    L = max(L0, F + 1)
    T = L - 1
    D, A = observations.shape[1], actions.shape[1]
    o, a, n, b = np.zeros((T, F * D)), np.zeros((T, F * A)), np.zeros((T, F * D)), np.zeros((T, F))
    for t in range(T):
        for i in range(F):
            if t + i < L0:
                o[t, i * D:(i + 1) * D] = observations[t + i]
                a[t, i * A:(i + 1) * A] = actions[t + i]
            if t + i + 1 < L0:
                n[t, i * D:(i + 1) * D] = observations[t + i + 1]
                b[t, i] = 1.0 if t > 0 else 0.0
    return o, a, n, b
