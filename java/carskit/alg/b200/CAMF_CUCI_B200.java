// CAMF_CUCI_B200.java -- CAMF_CUCI (src/carskit/alg/cars/adaptation/dependent/dev/CAMF_CUCI.java) on the B200 engine.
// The reference keeps icBias / ucBias in Guava HashBasedTables; initModel() fills EVERY (item | user, condition) cell
// (CAMF_CUCI.java:58-64), so the tables flatten to dense [numItems x C] / [numUsers x C] arrays without loss.
//     case "camf_cuci_b200": return new CAMF_CUCI_B200(trainMatrix, testMatrix, fold);
package carskit.alg.b200;

import java.util.ArrayList;
import java.util.List;

import com.google.common.collect.Table;

import carskit.alg.cars.adaptation.dependent.dev.CAMF_CUCI;
import carskit.b200.B200;
import carskit.b200.Native;
import carskit.data.structure.SparseMatrix;

public class CAMF_CUCI_B200 extends CAMF_CUCI {
    public CAMF_CUCI_B200(SparseMatrix trainMatrix, SparseMatrix testMatrix, int fold) {
        super(trainMatrix, testMatrix, fold);
        this.algoName = "CAMF_CUCI_B200";
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

    private static double[] flatten(Table<Integer, Integer, Double> t, int rows, int cols) {
        double[] out = new double[rows * cols];
        for (int i = 0; i < rows; i++)
            for (int c = 0; c < cols; c++)
                out[i * cols + c] = t.get(i, c);
        return out;
    }

    private static void unflatten(double[] flat, Table<Integer, Integer, Double> t, int rows, int cols) {
        for (int i = 0; i < rows; i++)
            for (int c = 0; c < cols; c++)
                t.put(i, c, flat[i * cols + c]);
    }

    /** Replaces the per-rating loop of CAMF_CUCI.buildModel() (CAMF_CUCI.java:78-134). */
    @Override
    protected void buildModel() throws Exception {
        B200.Ratings x = B200.flattenContextual(trainMatrix, rateDao);
        List<List<Integer>> conds = new ArrayList<>();
        for (int c = 0; c < rateDao.numContexts(); c++)
            conds.add(getConditions(c));
        int[][] ctx = B200.contextTable(conds);
        double[] fP = B200.flatten(P), fQ = B200.flatten(Q);
        double[] fIc = flatten(icBias, numItems, numConditions), fUc = flatten(ucBias, numUsers, numConditions);
        int mode = algoOptions != null && "fast".equalsIgnoreCase(algoOptions.getString("-mode", "exact")) ? Native.FAST : Native.EXACT;
        B200.train(Native.CAMF_CUCI, mode, numUsers, numItems, numConditions, numFactors, x, ctx, globalMean,
                (double) regU, (double) regI, (double) regB, (double) regC,
                B200.devicesFor(fold, algoOptions == null ? 1 : algoOptions.getInt("-gpus", 1)), numIters, control,
                fP, fQ, null, null, null, fIc, fUc, null, null);
        B200.unflatten(fP, P);
        B200.unflatten(fQ, Q);
        unflatten(fIc, icBias, numItems, numConditions);
        unflatten(fUc, ucBias, numUsers, numConditions);
    }
}
