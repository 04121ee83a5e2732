//go:build b200 && cgo

package cf

// #include "gorse_b200.h"
import "C"

import (
	"context"

	"github.com/gorse-io/gorse/common/log"
	"github.com/gorse-io/gorse/dataset"
	"go.uber.org/zap"
)

// Fit replaces (*BPR).Fit (model/cf/model.go:408-530).  In the reference file the original method is renamed
// fitCPU behind `//go:build !b200` (INTEGRATION.md); everything else on the type (SetParams, SuggestParams,
// Predict, Marshal, ...) is inherited unchanged, so master and worker are byte-identical.
func (bpr *BPR) Fit(ctx context.Context, trainSet, valSet dataset.CFSplit, config *FitConfig) Score {
	log.Logger().Info("fit bpr (b200)",
		zap.Int("train_set_size", trainSet.CountFeedback()),
		zap.Int("test_set_size", valSet.CountFeedback()),
		zap.Any("params", bpr.GetParams()),
		zap.Any("config", config))
	var p C.gorse_b200_fit_params
	C.gorse_b200_fit_params_default(0, &p)
	p.n_factors, p.n_epochs = C.int32_t(bpr.nFactors), C.int32_t(bpr.nEpochs)
	p.lr, p.reg = C.float(bpr.lr), C.float(bpr.reg)
	p.init_mean, p.init_stddev = C.float(bpr.initMean), C.float(bpr.initStdDev)
	p.seed = C.uint64_t(bpr.GetRandomGenerator().Int63()) // the model's RandomState stream, model/model.go:43
	return fitB200(ctx, &bpr.BaseMatrixFactorization, false, p, trainSet, valSet, config, "BPR.Fit")
}
