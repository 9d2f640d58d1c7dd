// B200.java -- what every *_B200 recommender shares: flattening the Java containers ONCE per buildModel() and the
// epoch loop around Native.epoch().  isConverged()/updateLRate() (IterativeRecommender.java:145-229) stay in the
// recommender: the loop hands every epoch's loss to EpochControl and reads the learning rate back.
package carskit.b200;

import java.util.ArrayList;
import java.util.List;

import carskit.data.processor.DataDAO;
import librec.data.DenseMatrix;
import librec.data.DenseVector;
import librec.data.SparseMatrix;

public final class B200 {
    private B200() {
    }

    /** The part of the iteration that stays in Java. */
    public interface EpochControl {
        /** lRate of the coming epoch (IterativeRecommender.lRate). */
        double lRate();

        /** Store the loss, run isConverged(iter); true = stop. */
        boolean afterEpoch(int iter, double loss) throws Exception;
    }

    /** Training ratings in the order `for (MatrixEntry me : trainMatrix)` yields them (CRS: row ascending, column ascending). */
    public static final class Ratings {
        public int[] u, j, ctx;
        public double[] r;
    }

    /** Context-aware models: rows of trainMatrix are user-item pair ids, columns are context ids (DataDAO.java:336-345). */
    public static Ratings flattenContextual(SparseMatrix trainMatrix, DataDAO rateDao) {
        int[] rp = trainMatrix.getRowPointers(), ci = trainMatrix.getColumnIndices();
        double[] rv = trainMatrix.getData();
        int nnz = rv.length;
        Ratings x = new Ratings();
        x.u = new int[nnz];
        x.j = new int[nnz];
        x.ctx = new int[nnz];
        x.r = new double[nnz];
        int n = 0;
        for (int ui = 0; ui + 1 < rp.length; ui++) {
            if (rp[ui] == rp[ui + 1])
                continue; // the iterator skips empty rows (SparseMatrix$MatrixIterator.nextNonEmptyRow)
            int uu = rateDao.getUserIdFromUI(ui), jj = rateDao.getItemIdFromUI(ui); // DataDAO.java:1038-1046
            for (int k = rp[ui]; k < rp[ui + 1]; k++, n++) {
                x.u[n] = uu;
                x.j[n] = jj;
                x.ctx[n] = ci[k];
                x.r[n] = rv[k];
            }
        }
        return x;
    }

    /** PMF / BiasedMF iterate the 2-D `train` (rows = users, columns = items; Recommender.java:252). */
    public static Ratings flatten2D(SparseMatrix train) {
        int[] rp = train.getRowPointers(), ci = train.getColumnIndices();
        double[] rv = train.getData();
        Ratings x = new Ratings();
        x.u = new int[rv.length];
        x.j = new int[rv.length];
        x.r = new double[rv.length];
        for (int uu = 0, n = 0; uu + 1 < rp.length; uu++)
            for (int k = rp[uu]; k < rp[uu + 1]; k++, n++) {
                x.u[n] = uu;
                x.j[n] = ci[k];
                x.r[n] = rv[k];
            }
        return x;
    }

    /** ctx -> condition ids in CSR form; conds.get(c) = getConditions(c) of the recommender (ContextRecommender.java:53-61). */
    public static int[][] contextTable(List<List<Integer>> conds) {
        int[] ptr = new int[conds.size() + 1];
        List<Integer> flat = new ArrayList<>();
        for (int c = 0; c < conds.size(); c++) {
            flat.addAll(conds.get(c));
            ptr[c + 1] = flat.size();
        }
        int[] cond = new int[flat.size()];
        for (int k = 0; k < cond.length; k++)
            cond[k] = flat.get(k);
        return new int[][] { ptr, cond };
    }

    public static double[] flatten(DenseMatrix m) {
        int rows = m.numRows(), cols = m.numColumns();
        double[] out = new double[rows * cols];
        for (int i = 0; i < rows; i++)
            for (int f = 0; f < cols; f++)
                out[i * cols + f] = m.get(i, f);
        return out;
    }

    public static void unflatten(double[] flat, DenseMatrix m) {
        int rows = m.numRows(), cols = m.numColumns();
        for (int i = 0; i < rows; i++)
            for (int f = 0; f < cols; f++)
                m.set(i, f, flat[i * cols + f]);
    }

    /** DenseVector.getData() returns the live double[] (bytecode: aload_0, getfield data, areturn); `size` is a protected field. */
    public static double[] flatten(DenseVector v) {
        return v.getData().clone();
    }

    public static void unflatten(double[] flat, DenseVector v) {
        System.arraycopy(flat, 0, v.getData(), 0, flat.length);
    }

    /** Which GPUs a recommender of cross-validation fold `fold` uses: -gpus in the algorithm's option line, else one, round-robin. */
    public static int[] devicesFor(int fold, int requested) {
        int have = Native.deviceCount();
        if (have <= 0)
            throw new RuntimeException("carskit_b200: no CUDA device; the engine has no CPU path");
        if (requested > 1) {
            int[] ids = new int[Math.min(requested, have)];
            for (int g = 0; g < ids.length; g++)
                ids[g] = g;
            return ids;
        }
        return new int[] { fold > 0 ? (fold - 1) % have : 0 };
    }

    /**
     * create + upload, then one Native.epoch per iteration with the Java-side convergence control in between, then
     * download into the SAME flat arrays (the caller unflattens them into P, Q, ...). Arrays the model lacks are null.
     */
    public static void train(int model, int mode, int numUsers, int numItems, int numConditions, int numFactors, Ratings x,
                             int[][] ctxTable, double globalMean, double regU, double regI, double regB, double regC,
                             int[] gpuIds, int numIters, EpochControl ctl, double[] P, double[] Q, double[] userBias,
                             double[] itemBias, double[] condBias, double[] icBias, double[] ucBias, double[] ccSim,
                             int[] emptyConditions) throws Exception {
        train(model, mode, numUsers, numItems, numConditions, numFactors, x, ctxTable, globalMean, regU, regI, regB, regC, gpuIds,
                numIters, ctl, P, Q, userBias, itemBias, condBias, icBias, ucBias, ccSim, emptyConditions, 0);
    }

    /** As above; `ccSim` is the similarity model's own array, numContextFactors CAMF_LCS's `-f`. */
    public static void train(int model, int mode, int numUsers, int numItems, int numConditions, int numFactors, Ratings x,
                             int[][] ctxTable, double globalMean, double regU, double regI, double regB, double regC,
                             int[] gpuIds, int numIters, EpochControl ctl, double[] P, double[] Q, double[] userBias,
                             double[] itemBias, double[] condBias, double[] icBias, double[] ucBias, double[] ccSim,
                             int[] emptyConditions, int numContextFactors) throws Exception {
        int numContexts = ctxTable == null ? 0 : ctxTable[0].length - 1;
        long h = Native.create(model, mode, numUsers, numItems, numConditions, numContexts, numFactors, x.u, x.j, x.ctx, x.r,
                ctxTable == null ? null : ctxTable[0], ctxTable == null ? null : ctxTable[1], globalMean, regU, regI, regB,
                regC, gpuIds, Native.COMBINE_MEAN, 0.0, emptyConditions, numContextFactors);
        try {
            Native.upload(h, P, Q, userBias, itemBias, condBias, icBias, ucBias, ccSim);
            for (int iter = 1; iter <= numIters; iter++) {
                double loss = Native.epoch(h, ctl.lRate()); // == the body of `for (int iter ...)` up to `loss *= 0.5`
                if (ctl.afterEpoch(iter, loss))
                    break;
            }
            Native.download(h, P, Q, userBias, itemBias, condBias, icBias, ucBias, ccSim);
        } finally {
            Native.destroy(h);
        }
    }
}
