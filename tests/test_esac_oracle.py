"""CPU: the torch restatement of EmotionSAC (oracle/nets_ref.py: esac_*) against fixtures recorded from the UNMODIFIED
reference agent (DifferentModules/ERL_Fill_agent.py, tests/golden/esac_golden.npz made by oracle/gen_golden.py esac)."""
import os

import numpy as np
import torch

from oracle import nets_ref as R

GOLD = os.path.join(os.path.dirname(__file__), "golden", "esac_golden.npz")


def _sd(G, prefix):
    return {k[len(prefix):]: torch.from_numpy(np.array(G[k])) for k in G.files if k.startswith(prefix)}


def _batch(G, idx):
    t = lambda k, dt=torch.float32: torch.from_numpy(np.asarray(G["buf_" + k])[idx]).to(dt)
    return {"states": t("states"), "actions": t("actions"), "rewards": t("rewards"), "next_states": t("next_states"),
            "dones": t("dones"), "emotions": t("emotions")}


def test_key_order_matches_reference_modules():
    G = np.load(GOLD)
    assert [k[3:] for k in G.files if k.startswith("p0.")] == R.ESAC_POLICY_ORDER
    assert [k[5:] for k in G.files if k.startswith("q1_0.")] == R.ESAC_Q_ORDER


def test_esac_update_matches_reference():
    G = np.load(GOLD)
    pol, q1, q2 = _sd(G, "p0."), _sd(G, "q1_0."), _sd(G, "q2_0.")
    q1t, q2t = {k: v.clone() for k, v in q1.items()}, {k: v.clone() for k, v in q2.items()}
    z = lambda w, order: [torch.zeros_like(w[k]) for k in order]
    opt = {"pm": z(pol, R.ESAC_POLICY_ORDER), "pv": z(pol, R.ESAC_POLICY_ORDER), "q1m": z(q1, R.ESAC_Q_ORDER),
           "q1v": z(q1, R.ESAC_Q_ORDER), "q2m": z(q2, R.ESAC_Q_ORDER), "q2v": z(q2, R.ESAC_Q_ORDER), "step": 0}
    e_prev = G["e_now"][0]
    for it in range(3):
        e_now = G["e_now"][it]
        res = R.esac_update(pol, q1, q2, q1t, q2t, opt, _batch(G, G["idx"][it]), torch.from_numpy(G["eps_next"][it]),
                            torch.from_numpy(G["eps_new"][it]), e_now, e_prev)
        e_prev = e_now
        np.testing.assert_allclose(res[:3], G["res"][it][:3], rtol=2e-5)
        assert abs(res[3] - G["res"][it][3]) <= 1e-6
    for name, w in (("p3.", pol), ("q1_3.", q1), ("q2_3.", q2), ("q1t_3.", q1t), ("q2t_3.", q2t)):
        for k, v in _sd(G, name).items():
            np.testing.assert_allclose(w[k].reshape(-1)[::9].numpy(), v.numpy(), rtol=2e-4, atol=2e-6, err_msg=name + k)


def test_esac_act_matches_reference():
    G = np.load(GOLD)
    pol = _sd(G, "p0.")      # act is recorded after the updates: use the end policy where the fixture holds all of it
    # the act fixture was taken with the END policy; only every 9th weight of it is stored, so pin act through the update:
    q1, q2 = _sd(G, "q1_0."), _sd(G, "q2_0.")
    q1t, q2t = {k: v.clone() for k, v in q1.items()}, {k: v.clone() for k, v in q2.items()}
    z = lambda w, order: [torch.zeros_like(w[k]) for k in order]
    opt = {"pm": z(pol, R.ESAC_POLICY_ORDER), "pv": z(pol, R.ESAC_POLICY_ORDER), "q1m": z(q1, R.ESAC_Q_ORDER),
           "q1v": z(q1, R.ESAC_Q_ORDER), "q2m": z(q2, R.ESAC_Q_ORDER), "q2v": z(q2, R.ESAC_Q_ORDER), "step": 0}
    e_prev = G["e_now"][0]
    for it in range(3):
        R.esac_update(pol, q1, q2, q1t, q2t, opt, _batch(G, G["idx"][it]), torch.from_numpy(G["eps_next"][it]),
                      torch.from_numpy(G["eps_new"][it]), G["e_now"][it], e_prev)
        e_prev = G["e_now"][it]
    lo = np.array([7000, 0, 24900, 35, 3, 5, 100, 100, 100, 300], np.float64)
    hi = np.array([18000, 1, 25000, 75, 5, 20, 300, 200, 200, 500], np.float64)
    off, sc = (hi + lo) / 2, (hi - lo) / 2
    e = torch.from_numpy(G["act_emotion"]).float().unsqueeze(0)
    for i in range(6):
        s = torch.from_numpy(G["act_states"][i]).unsqueeze(0)
        a, _ = R.esac_policy_sample(pol, s, e, torch.from_numpy(G["act_eps"][i]).unsqueeze(0))
        np.testing.assert_allclose(a[0].numpy().astype(np.float64) * sc + off, G["act_out"][i], rtol=1e-5, atol=1e-5 * sc.max())
        mean, _ = R.esac_policy_forward(pol, s, e)
        np.testing.assert_allclose(torch.tanh(mean)[0].numpy().astype(np.float64) * sc + off, G["act_det"][i], rtol=1e-5, atol=1e-5 * sc.max())
