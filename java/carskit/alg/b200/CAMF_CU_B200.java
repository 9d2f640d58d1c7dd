// CAMF_CU_B200.java -- CAMF_CU (src/carskit/alg/cars/adaptation/dependent/dev/CAMF_CU.java) with buildModel() on the B200 engine.
// Same constructor as the reference class; only buildModel() is overridden: initModel(), predict(), evalRatings(),
// evalRankings(), saveModel() are inherited and read the arrays this method writes back.
// Register beside the reference's own case in CARSKit.getRecommender (src/carskit/main/CARSKit.java:429-705):
//     case "camf_cu_b200": return new CAMF_CU_B200(trainMatrix, testMatrix, fold);
// Options (setting.conf, the algorithm's own line, e.g. `CAMF_CU_B200=-mode fast -gpus 8`): -mode exact|fast, -gpus N.
package carskit.alg.b200;

import java.util.ArrayList;
import java.util.List;

import carskit.alg.cars.adaptation.dependent.dev.CAMF_CU;
import carskit.b200.B200;
import carskit.b200.Native;
import carskit.data.structure.SparseMatrix;

public class CAMF_CU_B200 extends CAMF_CU {
    public CAMF_CU_B200(SparseMatrix trainMatrix, SparseMatrix testMatrix, int fold) {
        super(trainMatrix, testMatrix, fold);
        this.algoName = "CAMF_CU_B200";
    }

    private final B200.EpochControl control = new B200.EpochControl() {
        public double lRate() {
            return lRate;
        }

        public boolean afterEpoch(int iter, double epochLoss) throws Exception {
            loss = epochLoss;           // NaN / Inf included: isConverged() logs and exits (IterativeRecommender.java:181-184)
            return isConverged(iter);   // bold driver / decay / early stop, unchanged (IterativeRecommender.java:145-229)
        }
    };

    private int mode() {
        return algoOptions != null && "fast".equalsIgnoreCase(algoOptions.getString("-mode", "exact")) ? Native.FAST : Native.EXACT;
    }

    private int[] devices() {
        return B200.devicesFor(fold, algoOptions == null ? 1 : algoOptions.getInt("-gpus", 1));
    }

    private int[][] contextTable() {
        List<List<Integer>> conds = new ArrayList<>();
        for (int c = 0; c < rateDao.numContexts(); c++)
            conds.add(getConditions(c)); // ContextRecommender.java:53-61: the order predict() / buildModel() iterate
        return B200.contextTable(conds);
    }

    /** Replaces the per-rating loop of CAMF_CU.buildModel() (CAMF_CU.java:71-128). */
    @Override
    protected void buildModel() throws Exception {
        B200.Ratings x = B200.flattenContextual(trainMatrix, rateDao);
        int[][] ctx = contextTable();
        double[] fP = B200.flatten(P);
        double[] fQ = B200.flatten(Q);
        double[] fItemBias = B200.flatten(itemBias);
        double[] fUcBias = B200.flatten(ucBias);
        // float -> double widening of the static hyper-parameters (IterativeRecommender.java:40): never re-parse "0.001"
        B200.train(Native.CAMF_CU, mode(), numUsers, numItems, numConditions, numFactors, x, ctx, globalMean,
                (double) regU, (double) regI, (double) regB, (double) regC, devices(), numIters, control,
                fP, fQ, null, fItemBias, null, null, fUcBias, null, null);
        B200.unflatten(fP, P);
        B200.unflatten(fQ, Q);
        B200.unflatten(fItemBias, itemBias);
        B200.unflatten(fUcBias, ucBias);
    }
}
