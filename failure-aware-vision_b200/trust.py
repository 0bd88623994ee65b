"""f4 -- batched replay of the reference's TrustEngine (platform/backend/trust_engine.py:139-243) over many independent
tick sequences, the data-parallel form of the playground's replay loop (platform/backend/main.py:340-352).

    tr = fav.TrustReplay()
    res = tr.run(status, score, dt=1/30)        # status int [S, L] (STATUS codes), score float64 [S, L] (NaN = None)
    tr.state_dict(res, s, i)                    # TrustEngine.get_state()'s replay-determined keys, same rounding

There is no CPU fallback: the CUDA library must be present (oracle/trust.py is test infrastructure only)."""
import ctypes as C

import numpy as np
import torch

from . import _lib

STATUS = ("VISION_OK", "VISION_FROZEN", "VISION_BLANK", "VISION_CORRUPTED")
POLICY = ("VISION_ALLOWED", "VISION_DECLINING", "VISION_DEGRADED", "VISION_BLOCKED")
DECAY_RATES = {"VISION_OK": -0.10, "VISION_FROZEN": 0.30, "VISION_BLANK": 0.60, "VISION_CORRUPTED": 1.00}
FIELDS = ("reliability", "anomaly_integral", "trust_velocity", "recovery_debt", "recovery_coeff")


def status_codes(names):
    """Nested lists / array of VISION_* strings -> int8 codes."""
    lut = {n: i for i, n in enumerate(STATUS)}
    return np.vectorize(lambda n: lut[n], otypes=[np.int8])(np.asarray(names, dtype=object))


class TrustReplay:
    def __init__(self, handle=None, device=0):
        if not torch.cuda.is_available():
            raise RuntimeError("TrustReplay needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device("cuda", device)
        self.handle = handle if handle is not None else _lib.Handle(device)
        self.lib = self.handle.lib

    def run(self, status, score, dt=1.0 / 30.0, trajectory=True):
        """status [S, L] integer codes, score [S, L] float (NaN where the reference would pass None), dt a float or [L].
        Returns numpy arrays: 'final' [S, 8] and, with trajectory=True, 'state' [S, L, 5], 'policy', 'contradiction',
        'contradiction_count' [S, L]."""
        status = np.asarray(status)
        if status.size and (status.min() < 0 or status.max() > 3):
            # the reference engine only knows its four VISION_* strings (trust_engine.py:21-26); any other code would
            # decay at an undefined rate inside the kernel
            raise ValueError("status codes must be 0 (OK), 1 (FROZEN), 2 (BLANK) or 3 (CORRUPTED)")
        st = torch.as_tensor(status, dtype=torch.int8)
        sc = torch.as_tensor(np.asarray(score, dtype=np.float64))
        if st.dim() != 2 or sc.shape != st.shape:
            raise ValueError("status and score must both be [S, L]")
        S, L = st.shape
        dev = self.device
        st_d = st.to(dev).t().contiguous()                     # tick-major [L, S]
        sc_d = sc.to(dev).t().contiguous()
        dts = None
        if not np.isscalar(dt):
            dts = torch.as_tensor(np.asarray(dt, dtype=np.float64)).to(dev).contiguous()
            if dts.shape != (L,):
                raise ValueError("dt must be a scalar or have one entry per tick")
        final = torch.empty((S, 8), dtype=torch.float64, device=dev)
        state = policy = contra = count = None
        if trajectory:
            state = torch.empty((L, S, 5), dtype=torch.float64, device=dev)
            policy = torch.empty((L, S), dtype=torch.uint8, device=dev)
            contra = torch.empty((L, S), dtype=torch.uint8, device=dev)
            count = torch.empty((L, S), dtype=torch.int32, device=dev)
        p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        rc = self.lib.fav_trust_replay(self.handle.h, p(st_d), p(sc_d), p(dts), C.c_double(0.0 if dts is not None else float(dt)),
                                       S, L, p(state), p(policy), p(contra), p(count), p(final), stream)
        _lib.check(rc, "fav_trust_replay")
        out = {"final": final.cpu().numpy()}
        if trajectory:
            out.update(state=state.permute(1, 0, 2).contiguous().cpu().numpy(), policy=policy.t().contiguous().cpu().numpy(),
                       contradiction=contra.t().contiguous().cpu().numpy(), contradiction_count=count.t().contiguous().cpu().numpy())
        return out

    @staticmethod
    def state_dict(res, s, i, vision_status=None):
        """The keys of TrustEngine.get_state() (trust_engine.py:245-263) that the replay determines, with its rounding."""
        rel, integ, vel, debt, coeff = (float(v) for v in res["state"][s, i])
        d = {"reliability": round(rel, 6), "policy_state": POLICY[int(res["policy"][s, i])],
             "anomaly_integral": round(integ, 6), "trust_velocity": round(vel, 6), "recovery_debt": round(debt, 4),
             "recovery_coeff": round(coeff, 4), "contradiction_detected": bool(res["contradiction"][s, i]),
             "contradiction_count": int(res["contradiction_count"][s, i]), "recovery_coefficient": round(coeff, 4),
             "tick_count": i + 1}
        if vision_status is not None:
            d.update(vision_status=vision_status, ml_influence_active=vision_status == "VISION_OK",
                     decay_coefficient=DECAY_RATES.get(vision_status, 0))
        return d
