// Copyright 2026 gorse-b200 authors. Drop-in for gorse-io/gorse @ 5404aefa, package main of cmd/gorse-bench.
//
// NOT COMPILED IN THIS REPOSITORY (no Go toolchain in the build image).  The same sub-command exists as a C++ binary over the
// same C ABI, gorse_b200/csrc/gorse_bench_cf.cpp, which tests/test_ncf.py runs on the GPU box.
//
// `gorse-bench cf`: the collaborative-filtering sub-command cmd/gorse-bench lacks in this snapshot (it has `reranker` and
// `embedding`, main.go:61-64,558-559).  It loads a built-in NCF dataset with dataset.LoadDataFromBuiltIn
// (dataset/dataset.go:398-418), fits cf.BPR or cf.ALS -- on the GPU when built with -tags "b200 cgo", else the stock CPU Fit --
// and prints the tables the other sub-commands print (main.go:521-555).

//go:build cgo

package main

import (
	"context"
	"fmt"
	"os"
	"strconv"
	"time"

	"github.com/gorse-io/gorse/common/log"
	"github.com/gorse-io/gorse/dataset"
	"github.com/gorse-io/gorse/model"
	"github.com/gorse-io/gorse/model/cf"
	"github.com/olekukonko/tablewriter"
	"github.com/samber/lo"
	"github.com/spf13/cobra"
	"go.uber.org/zap"
)

var benchCFCmd = &cobra.Command{
	Use:   "cf",
	Short: "Benchmark collaborative filtering models (BPR, ALS)",
	Run: func(cmd *cobra.Command, args []string) {
		name, _ := cmd.Flags().GetString("dataset")
		modelName, _ := cmd.Flags().GetString("model")
		topK, _ := cmd.Flags().GetInt("top")
		jobs, _ := cmd.Flags().GetInt("jobs")
		train, test, err := dataset.LoadDataFromBuiltIn(name)
		if err != nil {
			log.Logger().Fatal("failed to load dataset", zap.Error(err))
		}
		table := tablewriter.NewWriter(os.Stdout)
		table.Header([]string{"", "#users", "#items", "#interactions"})
		lo.Must0(table.Bulk([][]string{
			{"train", strconv.Itoa(train.CountUsers()), strconv.Itoa(train.CountItems()), strconv.Itoa(train.CountFeedback())},
			{"test", strconv.Itoa(test.CountUsers()), strconv.Itoa(test.CountItems()), strconv.Itoa(test.CountFeedback())},
		}))
		lo.Must0(table.Render())

		params := model.Params{}
		if v, _ := cmd.Flags().GetInt("factors"); v > 0 {
			params[model.NFactors] = v
		}
		if v, _ := cmd.Flags().GetInt("epochs"); v > 0 {
			params[model.NEpochs] = v
		}
		var m cf.MatrixFactorization
		if modelName == "als" {
			m = cf.NewALS(params)
		} else {
			m = cf.NewBPR(params)
		}
		config := cf.NewFitConfig().SetJobs(jobs).SetTopK(topK)
		start := time.Now()
		score := m.Fit(context.Background(), train, test, config) // the b200 build routes this to gorse_b200_{bpr,als}_fit
		table = tablewriter.NewWriter(os.Stdout)
		table.Header([]string{"Model", fmt.Sprintf("NDCG@%d", topK), fmt.Sprintf("Precision@%d", topK), fmt.Sprintf("Recall@%d", topK), "sec"})
		lo.Must0(table.Bulk([][]string{{modelName, fmt.Sprintf("%.4f", score.NDCG), fmt.Sprintf("%.4f", score.Precision),
			fmt.Sprintf("%.4f", score.Recall), fmt.Sprintf("%.2f", time.Since(start).Seconds())}}))
		lo.Must0(table.Render())
	},
}

func init() {
	rootCmd.AddCommand(benchCFCmd)
	benchCFCmd.PersistentFlags().String("dataset", "ml-100k", "Built-in dataset in NCF format (ml-100k, ml-1m, pinterest-20)")
	benchCFCmd.PersistentFlags().String("model", "bpr", "bpr or als")
	benchCFCmd.PersistentFlags().Int("factors", 0, "Number of latent factors (0 = model default)")
	benchCFCmd.PersistentFlags().Int("epochs", 0, "Number of epochs (0 = model default)")
	benchCFCmd.PersistentFlags().IntP("top", "k", 10, "Number of top items to evaluate for each user")
}
