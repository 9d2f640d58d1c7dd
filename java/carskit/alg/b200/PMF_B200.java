// PMF_B200.java -- PMF (src/carskit/alg/baseline/cf/PMF.java) with buildModel() on the B200 engine.
// Same constructor as the reference class; only buildModel() is overridden: initModel(), predict(), evalRatings(),
// evalRankings(), saveModel() are inherited and read the arrays this method writes back.
// Register beside the reference's own case in CARSKit.getRecommender (src/carskit/main/CARSKit.java:429-705):
//     case "pmf_b200": return new PMF_B200(trainMatrix, testMatrix, fold);
// Options (setting.conf, the algorithm's own line, e.g. `PMF_B200=-mode fast -gpus 8`): -mode exact|fast, -gpus N.
package carskit.alg.b200;

import java.util.ArrayList;
import java.util.List;

import carskit.alg.baseline.cf.PMF;
import carskit.b200.B200;
import carskit.b200.Native;
import carskit.data.structure.SparseMatrix;

public class PMF_B200 extends PMF {
    public PMF_B200(SparseMatrix trainMatrix, SparseMatrix testMatrix, int fold) {
        super(trainMatrix, testMatrix, fold);
        this.algoName = "PMF_B200";
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

    /** Replaces the per-rating loop of PMF.buildModel() (PMF.java:47-82). */
    @Override
    protected void buildModel() throws Exception {
        B200.Ratings x = B200.flatten2D(train); // the 2-D `train` = rateDao.toTraditionalSparseMatrix(trainMatrix) (Recommender.java:252)
        int[][] ctx = null;
        double[] fP = B200.flatten(P);
        double[] fQ = B200.flatten(Q);
        // float -> double widening of the static hyper-parameters (IterativeRecommender.java:40): never re-parse "0.001"
        B200.train(Native.PMF, mode(), numUsers, numItems, 0, numFactors, x, ctx, globalMean,
                (double) regU, (double) regI, (double) regB, (double) regC, devices(), numIters, control,
                fP, fQ, null, null, null, null, null, null, null);
        B200.unflatten(fP, P);
        B200.unflatten(fQ, Q);
    }
}
