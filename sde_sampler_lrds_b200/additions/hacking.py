"""Evaluation helpers with the reference's names (sde_sampler/additions/hacking.py): ``evaluate_eubo`` (14-33) and
``TrainableWrapper.evaluate / compute_results_eubo`` (68-91).  The forward (EUBO) estimators come from the same fp64
partials as the backward ones (estimators.py); ``TrainableWrapper.run`` (36-66) trains with the solver's ``step``."""
from __future__ import annotations

import time

import torch

from ..estimators import estimator_partials, metrics_from_partials


def evaluate_eubo(trainable, results, compute_eubo_last_arg, use_ema, group=None):
    """log Z_f, EUBO and forward ESS from a noising rollout started at target samples (hacking.py:14-33)."""
    eval_samples_from_target = trainable.target.sample((trainable.eval_batch_size,))
    with torch.no_grad():
        rnd_target = trainable.loss.compute_eubo(trainable.eval_ts.to(trainable.device), eval_samples_from_target,
                                                 trainable.clipped_target_unnorm_log_prob, compute_eubo_last_arg,
                                                 use_ema=use_ema)
    m = metrics_from_partials(estimator_partials(rnd_target, group=group))
    n = m["count"]
    # weights = softmax(-rnd): 1 / sum w^2 = (sum e^{-rnd-m})^2 / sum e^{2(-rnd-m)}
    results.metrics["eval/log_norm_const_is_f"] = m["log_norm_const_is_f"]
    results.metrics["eval/eubo"] = m["eubo"]
    results.metrics["eval/effective_sample_size_f"] = m["effective_sample_size"]
    results.metrics["eval/norm_effective_sample_size_f"] = m["effective_sample_size"] / n
    return results


class TrainableWrapper(torch.nn.Module):
    def __init__(self, trainable, verbose=True):
        super().__init__()
        self.trainable = trainable
        self.verbose = verbose

    def run(self, keep_training_metrics=False):
        """train_steps stochastic gradient steps, then the evaluation with the EUBO metrics (hacking.py:43-66)."""
        t = self.trainable
        t.train()
        training_metrics, training_time = [], 0.0
        for i in range(t.n_steps, t.train_steps):
            start = time.time()
            metrics = t.step(i)
            training_time += time.time() - start
            if keep_training_metrics:
                training_metrics.append(metrics)
            if self.verbose:
                print("step {} loss={:.2e}".format(i, metrics["train/loss"]))
        results = self.evaluate(use_ema=t.use_ema)
        results.metrics["eval/training_time"] = training_time
        if keep_training_metrics:
            return results, {k: [m[k] for m in training_metrics if k in m] for k in training_metrics[0]}
        return results

    def compute_results_eubo(self, results, use_ema=True):
        t = self.trainable
        if hasattr(t.loss, "compute_eubo") and t.eubo_available and hasattr(t.target, "sample"):
            last = t.reference_distr.log_prob if hasattr(t, "reference_distr") else t.prior.log_prob
            results = evaluate_eubo(t, results, last, use_ema=use_ema)
        return results

    @torch.no_grad()
    def evaluate(self, use_ema=True, log=True):
        use_ema_ = self.trainable.use_ema and use_ema
        results = self.trainable.compute_results(use_ema=use_ema_)
        return self.compute_results_eubo(results, use_ema=use_ema_)
