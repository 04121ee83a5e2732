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

// Fit replaces (*ALS).Fit, the eALS / "CCD" path (model/cf/model.go:609-775).
func (als *ALS) Fit(ctx context.Context, trainSet, valSet dataset.CFSplit, config *FitConfig) Score {
	log.Logger().Info("fit als (b200)",
		zap.Int("train_set_size", trainSet.CountFeedback()),
		zap.Int("test_set_size", valSet.CountFeedback()),
		zap.Any("params", als.GetParams()),
		zap.Any("config", config))
	var p C.gorse_b200_fit_params
	C.gorse_b200_fit_params_default(1, &p)
	p.n_factors, p.n_epochs = C.int32_t(als.nFactors), C.int32_t(als.nEpochs)
	p.reg, p.alpha = C.float(als.reg), C.float(als.weight)
	p.init_mean, p.init_stddev = C.float(als.initMean), C.float(als.initStdDev)
	p.seed = C.uint64_t(als.GetRandomGenerator().Int63())
	return fitB200(ctx, &als.BaseMatrixFactorization, true, p, trainSet, valSet, config, "ALS.Fit")
}
