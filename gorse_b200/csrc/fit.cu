// fit.cu -- host-side mirror of cf.BPR.Fit / cf.ALS.Fit (model/cf/model.go:408-530, 609-775) above the
// per-epoch entry points: Init, Evaluate at epoch 0, epoch loop, Evaluate every `verbose` epochs and on the last
// one, early stopping on `patience`, progress / cancellation callback (monitor span.Add(1) and ctx.Err()).
// The Go shim can call this once per Fit, or drive the per-epoch entry points itself and keep the loop in Go.
#include <vector>

#include "cf.cuh"

using namespace gb;

namespace {

struct EvalSets {
    const int64_t *test_off;
    const int32_t *test_items;
    const int64_t *neg_off;
    const int32_t *neg_items;
};

int32_t fit_loop(gorse_b200_cf *cf, bool als, const gorse_b200_fit_params *p, const EvalSets &ev, gorse_b200_progress_fn progress,
                 void *user, gorse_b200_fit_result *res)
{
    float score[3] = {0.f, 0.f, 0.f};
    std::vector<std::pair<int32_t, float>> scores;  // (epoch, NDCG) -- model.go:434,499
    *res = gorse_b200_fit_result{};
    // Init (model.go:414 / :615)
    GB_TRY(gorse_b200_cf_init_normal(cf, p->init_mean, p->init_stddev, p->seed));
    // Evaluate's inputs go to the device once per Fit; without caller-sampled negatives they are drawn there
    // (testSet.SampleUserNegatives(trainSet, Candidates) with NewRandomGenerator(0), evaluator.go:42, dataset.go:244)
    gorse_b200_eval *plan = nullptr;
    GB_TRY(gorse_b200_eval_create(cf, ev.test_off, ev.test_items, ev.neg_off, ev.neg_items, p->candidates, 0, p->topk, &plan));
    struct PlanGuard { gorse_b200_eval *p; ~PlanGuard() { gorse_b200_eval_destroy(p); } } guard{plan};
    GB_TRY(gorse_b200_eval_run(plan, score));  // :433 / :630
    scores.emplace_back(0, score[0]);
    int32_t epochs_run = 0;
    for (int32_t epoch = 1; epoch <= p->n_epochs; epoch++) {
        if (als) GB_TRY(gorse_b200_als_epoch(cf, p->reg, p->alpha));
        else GB_TRY(gorse_b200_bpr_epoch(cf, p->lr, p->reg, cf->n_feedback_global, p->seed + (uint64_t)epoch, GORSE_B200_SCATTER_ATOMIC));
        epochs_run = epoch;
        bool evaluated = false;
        // cross validation cadence, model.go:496 / :741
        if ((p->verbose > 0 && epoch % p->verbose == 0) || epoch == p->n_epochs) {
            GB_TRY(gorse_b200_eval_run(plan, score));
            scores.emplace_back(epoch, score[0]);
            evaluated = true;
        }
        // span.Add(1) + ctx cancellation (model.go:490-493,519): a non-zero return cancels -> Score{}
        if (progress && progress(user, epoch, p->n_epochs, evaluated ? score[0] : -1.0f) != 0) {
            GB_TRY(gorse_b200_ctx_sync(cf->ctx));
            *res = gorse_b200_fit_result{};
            res->epochs_run = epochs_run;
            res->cancelled = 1;
            return GORSE_B200_OK;
        }
        // early stopping if no improvement in the last `patience` epochs, model.go:508-517
        if (evaluated && p->patience > 0 && epoch > p->patience) {
            // lo.MaxBy with a.B > b.B keeps the FIRST maximum
            std::pair<int32_t, float> best = scores[0];
            for (auto &s : scores) if (s.second > best.second) best = s;
            if (best.first <= epoch - p->patience) {
                res->early_stopped = 1;
                res->best_epoch = best.first;
                break;
            }
        }
    }
    GB_TRY(gorse_b200_ctx_sync(cf->ctx));
    res->ndcg = score[0];
    res->precision = score[1];
    res->recall = score[2];
    res->epochs_run = epochs_run;
    return GORSE_B200_OK;
}

int32_t check_params(const gorse_b200_fit_params *p, const gorse_b200_fit_result *res)
{
    GB_CHECK_ARG(p != nullptr && res != nullptr, "NULL params/result");
    GB_CHECK_ARG(p->n_epochs >= 0, "negative n_epochs");
    GB_CHECK_ARG(p->topk >= 1, "topk must be >= 1");
    GB_CHECK_ARG(p->candidates >= 0, "negative candidates");
    return GORSE_B200_OK;
}

}  // namespace

extern "C" {

int32_t gorse_b200_fit_params_default(int32_t als, gorse_b200_fit_params *p)
{
    GB_CHECK_ARG(p != nullptr, "params is NULL");
    *p = gorse_b200_fit_params{};
    // BPR.SetParams model.go:389-394 / ALS.SetParams :580-585; FitConfig NewFitConfig :58-65
    p->n_factors = 16;
    p->n_epochs = als ? 50 : 100;
    p->lr = 0.05f;
    p->reg = als ? 0.06f : 0.01f;
    p->init_mean = 0.f;
    p->init_stddev = als ? 0.1f : 0.001f;
    p->alpha = 0.001f;
    p->seed = 0;
    p->verbose = 10;
    p->candidates = 100;
    p->topk = 10;
    p->patience = 0;
    return GORSE_B200_OK;
}

int32_t gorse_b200_bpr_fit(gorse_b200_cf *cf, const gorse_b200_fit_params *params, const int64_t *test_off,
                           const int32_t *test_items, const int64_t *neg_off, const int32_t *neg_items,
                           gorse_b200_progress_fn progress, void *user, gorse_b200_fit_result *result)
{
    GB_CHECK_ARG(cf != nullptr, "cf is NULL");
    GB_TRY(check_params(params, result));
    EvalSets ev{test_off, test_items, neg_off, neg_items};
    return fit_loop(cf, false, params, ev, progress, user, result);
}

int32_t gorse_b200_als_fit(gorse_b200_cf *cf, const gorse_b200_fit_params *params, const int64_t *test_off,
                           const int32_t *test_items, const int64_t *neg_off, const int32_t *neg_items,
                           gorse_b200_progress_fn progress, void *user, gorse_b200_fit_result *result)
{
    GB_CHECK_ARG(cf != nullptr, "cf is NULL");
    GB_TRY(check_params(params, result));
    EvalSets ev{test_off, test_items, neg_off, neg_items};
    return fit_loop(cf, true, params, ev, progress, user, result);
}

}  // extern "C"
