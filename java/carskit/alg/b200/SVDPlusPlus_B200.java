// SVDPlusPlus_B200.java -- SVD++ (src/carskit/alg/baseline/cf/SVDPlusPlus.java) with buildModel() on the B200 engine.
// Every rating of user u rewrites Y[k] for all items k the user rated: the dependency DAG is one rating wide, so the
// engine runs it as one chain on one warp -- bit-identical to the reference, for the small data sets that allows.
//     case "svd++_b200": return new SVDPlusPlus_B200(trainMatrix, testMatrix, fold);
package carskit.alg.b200;

import carskit.alg.baseline.cf.SVDPlusPlus;
import carskit.b200.B200;
import carskit.b200.Native;
import carskit.data.structure.SparseMatrix;

public class SVDPlusPlus_B200 extends SVDPlusPlus {
    public SVDPlusPlus_B200(SparseMatrix trainMatrix, SparseMatrix testMatrix, int fold) {
        super(trainMatrix, testMatrix, fold);
        this.algoName = "SVD++_B200";
    }

    private final B200.EpochControl control = new B200.EpochControl() {
        public double lRate() {
            return lRate;
        }

        public boolean afterEpoch(int iter, double epochLoss) throws Exception {
            loss = epochLoss;
            return isConverged(iter);
        }
    };

    /** Replaces the per-rating loop of SVDPlusPlus.buildModel() (SVDPlusPlus.java:55-124); userItemsCache is rebuilt by the engine. */
    @Override
    protected void buildModel() throws Exception {
        B200.Ratings x = B200.flatten2D(train);
        double[] fP = B200.flatten(P), fQ = B200.flatten(Q), fY = B200.flatten(Y);
        double[] fUserBias = B200.flatten(userBias), fItemBias = B200.flatten(itemBias);
        B200.train(Native.SVDPP, Native.EXACT, numUsers, numItems, 0, numFactors, x, null, globalMean,
                (double) regU, (double) regI, (double) regB, (double) regC, B200.devicesFor(fold, 1), numIters, control,
                fP, fQ, fUserBias, fItemBias, null, null, null, fY, null);
        B200.unflatten(fP, P);
        B200.unflatten(fQ, Q);
        B200.unflatten(fY, Y);
        B200.unflatten(fUserBias, userBias);
        B200.unflatten(fItemBias, itemBias);
    }
}
