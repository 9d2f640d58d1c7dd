// CAMF_LCS_B200.java -- CAMF_LCS (src/carskit/alg/cars/adaptation/dependent/sim/CAMF_LCS.java) with buildModel() on the
// B200 engine (EXACT mode: one chain through cfMatrix_LCS, so one warp -- meant for the small data sets the model targets).
//     case "camf_lcs_b200": return new CAMF_LCS_B200(trainMatrix, testMatrix, fold);
package carskit.alg.b200;

import java.util.ArrayList;
import java.util.List;

import carskit.alg.cars.adaptation.dependent.sim.CAMF_LCS;
import carskit.b200.B200;
import carskit.b200.Native;
import carskit.data.structure.SparseMatrix;

public class CAMF_LCS_B200 extends CAMF_LCS {
    public CAMF_LCS_B200(SparseMatrix trainMatrix, SparseMatrix testMatrix, int fold) {
        super(trainMatrix, testMatrix, fold);
        this.algoName = "CAMF_LCS_B200";
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

    /** Replaces the per-rating loop of CAMF_LCS.buildModel() (CAMF_LCS.java:66-146). */
    @Override
    protected void buildModel() throws Exception {
        B200.Ratings x = B200.flattenContextual(trainMatrix, rateDao);
        List<List<Integer>> conds = new ArrayList<>();
        for (int c = 0; c < rateDao.numContexts(); c++)
            conds.add(getConditions(c));
        int[][] ctx = B200.contextTable(conds);
        int numF = cfMatrix_LCS.numColumns(); // `numF` itself is private to CAMF_LCS (:29); initModel() sized the matrix with it
        double[] fP = B200.flatten(P), fQ = B200.flatten(Q), cf = B200.flatten(cfMatrix_LCS);
        int[] empty = new int[EmptyContextConditions.size()];
        for (int i = 0; i < empty.length; i++)
            empty[i] = EmptyContextConditions.get(i);
        B200.train(Native.CAMF_LCS, Native.EXACT, numUsers, numItems, numConditions, numFactors, x, ctx, globalMean,
                (double) regU, (double) regI, (double) regB, (double) regC, B200.devicesFor(fold, 1), numIters, control,
                fP, fQ, null, null, null, null, null, cf, empty, numF);
        B200.unflatten(fP, P);
        B200.unflatten(fQ, Q);
        B200.unflatten(cf, cfMatrix_LCS);
    }
}
