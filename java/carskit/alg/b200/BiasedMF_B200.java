// BiasedMF_B200.java -- BiasedMF (src/carskit/alg/baseline/cf/BiasedMF.java) with buildModel() on the B200 engine.
// Same constructor as the reference class; only buildModel() is overridden: initModel(), predict(), evalRatings(),
// evalRankings(), saveModel() are inherited and read the arrays this method writes back.
// Register beside the reference's own case in CARSKit.getRecommender (src/carskit/main/CARSKit.java:429-705):
//     case "biasedmf_b200": return new BiasedMF_B200(trainMatrix, testMatrix, fold);
// Options (setting.conf, the algorithm's own line, e.g. `BiasedMF_B200=-mode fast -gpus 8`): -mode exact|fast, -gpus N.
package carskit.alg.b200;

import java.util.ArrayList;
import java.util.List;

import carskit.alg.baseline.cf.BiasedMF;
import carskit.b200.B200;
import carskit.b200.Native;
import carskit.data.structure.SparseMatrix;

public class BiasedMF_B200 extends BiasedMF {
    public BiasedMF_B200(SparseMatrix trainMatrix, SparseMatrix testMatrix, int fold) {
        super(trainMatrix, testMatrix, fold);
        this.algoName = "BiasedMF_B200";
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

    /** Replaces the per-rating loop of BiasedMF.buildModel() (BiasedMF.java:58-108). */
    @Override
    protected void buildModel() throws Exception {
        B200.Ratings x = B200.flatten2D(train); // the 2-D `train` = rateDao.toTraditionalSparseMatrix(trainMatrix) (Recommender.java:252)
        int[][] ctx = null;
        double[] fP = B200.flatten(P);
        double[] fQ = B200.flatten(Q);
        double[] fUserBias = B200.flatten(userBias);
        double[] fItemBias = B200.flatten(itemBias);
        // float -> double widening of the static hyper-parameters (IterativeRecommender.java:40): never re-parse "0.001"
        B200.train(Native.BIASEDMF, mode(), numUsers, numItems, 0, numFactors, x, ctx, globalMean,
                (double) regU, (double) regI, (double) regB, (double) regC, devices(), numIters, control,
                fP, fQ, fUserBias, fItemBias, null, null, null, null, null);
        B200.unflatten(fP, P);
        B200.unflatten(fQ, Q);
        B200.unflatten(fUserBias, userBias);
        B200.unflatten(fItemBias, itemBias);
    }
}
