(ns raytrace-clj.native
  "JNA binding + marshaller for libraytrace_b200.so (include/raytrace_b200.h).

  NOT executed in the build environment (no JVM there); shipped as the source a maintainer adds to
  gonewest818/raytrace-clj next to src/raytrace_clj/core.clj.  It replaces the render block
  core.clj:99-108 with one native call and leaves arg parsing (core.clj:74-80), scene building
  (core.clj:82-90) and saving (core.clj:112) untouched.

  project.clj: add [net.java.dev.jna/jna \"5.14.0\"] to :dependencies (project.clj:6-16)."
  (:require [clojure.core.matrix :as mat]
            [mikera.image.core]
            [mikera.image.colours]
            [raytrace-clj.perlin])
  (:import [com.sun.jna Native Pointer Memory Structure Function NativeLibrary]
           [com.sun.jna.ptr PointerByReference]
           [raytrace_clj.hitable Sphere UVSphere MovingSphere Hitlist bvh-node RectXY RectXZ RectYZ FlipNormals Translate RotateY
            Box ConstantMedium Triangle]
           [raytrace_clj.shader Lambertian Metal Dielectric DiffuseLight Isotropic]
           [raytrace_clj.texture Constant UVGradient Checkerboard PerlinNoise PerlinTurbulence Marble FlipTextureU FlipTextureV ImageMap]
           [raytrace_clj.camera ThinLensCamera PinholeCamera]))

(def ^:private lib (delay (NativeLibrary/getInstance "raytrace_b200")))
(defn- f ^Function [name] (.getFunction ^NativeLibrary @lib name))

(defn- check [ctx rc what]
  (when-not (zero? rc)
    (throw (ex-info (str what " failed: " (.invokeString (f "rt_last_error") (to-array [ctx]) false))
                    {:code rc}))))

(declare v3 floats->mem ints->mem)

;; ---- flatten the world: bvh-node tree (hitable.clj:97-123) -> leaves in left-to-right order.
;; A 1-element bvh-node stores the same object as both children (hitable.clj:113-114): visited once.  Wrappers
;; (FlipNormals / Translate / RotateY, hitable.clj:375-486) fold into a per-leaf op chain, outermost first; a Box
;; (hitable.clj:491-511) contributes its six rectangles.  Returns [[leaf ops] ...], ops = [[op p0 p1 p2 p3] ...].
(def ^:const XOP-TRANSLATE 1) (def ^:const XOP-ROTATE-Y 2) (def ^:const XOP-FLIP 3)
(defn flatten-world [world]
  (let [out (java.util.ArrayList.)]
    (letfn [(walk [h ops]
              (cond
                (instance? bvh-node h) (do (walk (:left h) ops) (when-not (identical? (:left h) (:right h)) (walk (:right h) ops)))
                (instance? Hitlist h) (run! #(walk % ops) (:items h))
                (instance? Box h) (walk (:sides h) ops)
                (instance? FlipNormals h) (walk (:item h) (conj ops [XOP-FLIP 0 0 0 0]))
                (instance? Translate h) (walk (:item h) (conj ops (into [XOP-TRANSLATE] (conj (v3 (:offset h)) 0))))
                (instance? RotateY h) (walk (:obj h) (conj ops [XOP-ROTATE-Y (:sin-theta h) (:cos-theta h) 0 0]))
                (or (instance? Sphere h) (instance? UVSphere h) (instance? MovingSphere h) (instance? RectXY h) (instance? RectXZ h)
                    (instance? RectYZ h) (instance? Triangle h) (instance? ConstantMedium h))
                (.add out [h ops])
                :else (throw (ex-info (str (type h) " is outside the accelerated path") {:code -2}))))]
      (walk world []))
    (vec out)))

(defn- floats->mem ^Memory [xs]
  (let [n (count xs) m (Memory. (max 4 (* 4 n)))]
    (.write m 0 (float-array xs) 0 n) m))
(defn- ints->mem ^Memory [xs]
  (let [n (count xs) m (Memory. (max 4 (* 4 n)))]
    (.write m 0 (int-array xs) 0 n) m))
(defn- v3 [v] [(mat/mget v 0) (mat/mget v 1) (mat/mget v 2)])

(defn- prim-row
  "one leaf -> the per-primitive fields of rt_scene_desc / rt_scene_ext"
  [s xform mat]
  (let [z12 (vec (repeat 12 0.0))
        base {:c0r [0 0 0 0] :c1 [0 0 0 0] :tt [0 1] :flags 0 :type 0 :q z12 :aux [0 0] :xform xform :mat mat}]
    (cond
      (instance? MovingSphere s) (assoc base :c0r (conj (v3 (:center0 s)) (:radius s)) :c1 (conj (v3 (:center1 s)) 0)
                                        :tt [(:t0 s) (:t1 s)] :flags 2)
      (or (instance? Sphere s) (instance? UVSphere s))
      (assoc base :c0r (conj (v3 (:center s)) (:radius s)) :c1 (conj (v3 (:center s)) 0) :flags (if (instance? UVSphere s) 1 0))
      (instance? RectXY s) (assoc base :type 1 :q (into [(:x0 s) (:y0 s) (:x1 s) (:y1 s) (:k s)] (repeat 7 0.0)))
      (instance? RectXZ s) (assoc base :type 2 :q (into [(:x0 s) (:z0 s) (:x1 s) (:z1 s) (:k s)] (repeat 7 0.0)))
      (instance? RectYZ s) (assoc base :type 3 :q (into [(:y0 s) (:z0 s) (:y1 s) (:z1 s) (:k s)] (repeat 7 0.0)))
      (instance? Triangle s) (assoc base :type 4 :q (into (vec (mapcat v3 [(:v0 s) (:v1 s) (:v2 s)])) (repeat 3 0.0))))))

(defn marshal-world
  "world -> map of JNA Memory blocks laid out as rt_scene_desc + rt_scene_ext expect (ABI 2)"
  [world]
  (let [leaves (flatten-world world)
        texs (java.util.ArrayList.) tex-ix (java.util.IdentityHashMap.)
        mats (java.util.ArrayList.) mat-ix (java.util.IdentityHashMap.)
        images (java.util.ArrayList.)
        xforms (java.util.ArrayList.) xform-ix (java.util.HashMap.)
        add-tex (fn add-tex [t]
                  (or (.get tex-ix t)
                      (let [z (fn [n] (repeat n 0))
                            rec (cond
                                  (instance? Constant t) {:type 0 :p (concat (v3 (:color t)) (z 9)) :ch [-1 -1]}
                                  (instance? UVGradient t) {:type 1 :p (mapcat v3 [(:co t) (:cu t) (:cv t) (:cuv t)]) :ch [-1 -1]}
                                  (instance? Checkerboard t) {:type 2 :p (cons (:scale t) (z 11))
                                                              :ch [(add-tex (:tex0 t)) (add-tex (:tex1 t))]}
                                  (instance? PerlinNoise t) {:type 3 :p (cons (:scale t) (z 11)) :ch [-1 -1]}
                                  (instance? PerlinTurbulence t) {:type 4 :p (concat [(:scale t) (:depth t)] (z 10)) :ch [-1 -1]}
                                  (instance? Marble t) {:type 5 :p (concat [(:scale t) (:depth t)] (z 10)) :ch [-1 -1]}
                                  (instance? FlipTextureU t) {:type 6 :p (z 12) :ch [(add-tex (:tex t)) -1]}
                                  (instance? FlipTextureV t) {:type 7 :p (z 12) :ch [(add-tex (:tex t)) -1]}
                                  (instance? ImageMap t) (let [i (.size images)] (.add images (:image t))
                                                           {:type 8 :p (cons i (z 11)) :ch [-1 -1]})
                                  :else (throw (ex-info (str "texture " (type t) " unsupported") {:code -2})))
                            i (.size texs)]
                        (.add texs rec) (.put tex-ix t i) i)))
        add-mat (fn [m]
                  (or (.get mat-ix m)
                      (let [rec (cond
                                  (instance? Lambertian m) [0 0.0 (add-tex (:albedo m))]
                                  (instance? Metal m) [1 (:fuzz m) (add-tex (:albedo m))]
                                  (instance? Dielectric m) [2 (:ri m) -1]
                                  (instance? DiffuseLight m) [3 0.0 (add-tex (:tex m))]
                                  (instance? Isotropic m) [4 0.0 (add-tex (:albedo m))]
                                  :else (throw (ex-info (str "material " (type m) " unsupported") {:code -2})))
                            i (.size mats)]
                        (.add mats rec) (.put mat-ix m i) i)))
        add-xform (fn [ops]
                    (cond (empty? ops) -1
                          (> (count ops) 4) (throw (ex-info "more than 4 nested wrappers around one leaf" {:code -2}))
                          :else (or (.get xform-ix ops) (let [i (.size xforms)] (.add xforms ops) (.put xform-ix ops i) i))))
        world-rows (java.util.ArrayList.) bnd-rows (java.util.ArrayList.) media (java.util.ArrayList.)
        n-world (count leaves)]
    (doseq [[s ops] leaves]
      (if (instance? ConstantMedium s)
        (let [row {:c0r [0 0 0 0] :c1 [0 0 0 0] :tt [0 1] :flags 0 :type 5 :q (into [(:density s)] (repeat 11 0.0)) :aux [0 0]
                   :xform (add-xform ops) :mat (add-mat (:phase-fn s))}]
          (.add media [(.size world-rows) (:boundary s)]) (.add world-rows row))
        (.add world-rows (prim-row s (add-xform ops) (add-mat (:material s))))))
    (doseq [[i bnd] media]
      (let [bl (flatten-world bnd) first-b (+ n-world (.size bnd-rows))]
        (doseq [[b bops] bl] (.add bnd-rows (prim-row b (add-xform bops) 0)))
        (.set world-rows i (assoc (.get world-rows i) :aux [first-b (count bl)]))))
    (let [rows (concat world-rows bnd-rows)
          xo (mapcat (fn [ops] (take 4 (concat (map first ops) (repeat 0)))) xforms)
          xp (mapcat (fn [ops] (take 16 (concat (mapcat rest ops) (repeat 0.0)))) xforms)
          img-bytes (map (fn [im] (let [w (mikera.image.core/width im) h (mikera.image.core/height im)]
                                    {:w w :h h :rgb (byte-array (for [y (range h) x (range w)
                                                                      c (take 3 (mikera.image.colours/components-rgb (mikera.image.core/get-pixel im x y)))]
                                                                  (unchecked-byte c)))})) images)
          offsets (reductions + 0 (map #(alength ^bytes (:rgb %)) img-bytes))
          rgb-mem (let [total (last offsets) m (Memory. (max 4 total))]
                    (doseq [[o im] (map vector offsets img-bytes)] (.write m (long o) ^bytes (:rgb im) 0 (alength ^bytes (:rgb im)))) m)
          uses-perlin (some #(#{3 4 5} (:type %)) texs)]
      {:n n-world :n-boundary (.size bnd-rows)
       :center0-r (floats->mem (mapcat :c0r rows)) :center1 (floats->mem (mapcat :c1 rows))
       :t0t1 (floats->mem (mapcat :tt rows)) :flags (ints->mem (map :flags rows)) :mat-id (ints->mem (map :mat rows))
       :n-mat (.size mats) :mat-type (ints->mem (map first mats)) :mat-param (floats->mem (map second mats))
       :mat-tex (ints->mem (map #(nth % 2) mats))
       :n-tex (.size texs) :tex-type (ints->mem (map :type texs)) :tex-params (floats->mem (mapcat :p texs))
       :tex-children (ints->mem (mapcat :ch texs))
       :prim-type (ints->mem (map :type rows)) :prim-params (floats->mem (mapcat :q rows)) :prim-aux (ints->mem (mapcat :aux rows))
       :prim-xform (ints->mem (map :xform rows)) :n-xforms (.size xforms) :xform-ops (ints->mem xo) :xform-params (floats->mem xp)
       :tie-rule (if (instance? bvh-node world) 1 0)
       ;; perlin.clj:6-17: the tables of THIS JVM (namespace-level defs) travel with the scene
       :perlin-vectors (when uses-perlin (floats->mem (mapcat v3 raytrace-clj.perlin/random-vectors)))
       :perlin-perm (when uses-perlin (ints->mem (concat raytrace-clj.perlin/perm-x raytrace-clj.perlin/perm-y raytrace-clj.perlin/perm-z)))
       :n-images (count img-bytes) :image-wh (ints->mem (mapcat (juxt :w :h) img-bytes))
       :image-offset (let [m (Memory. (max 8 (* 8 (count img-bytes))))] (doseq [[i o] (map-indexed vector (butlast offsets))] (.setLong m (* 8 i) o)) m)
       :image-rgb rgb-mem})))

(defn- scene-desc ^Memory [{:keys [n center0-r center1 t0t1 flags mat-id n-mat mat-type mat-param mat-tex
                                   n-tex tex-type tex-params tex-children]}]
  ;; struct rt_scene_desc on LP64: int32 + pad, 5 pointers, int32 + pad, 3 pointers, int32 + pad, 3 pointers
  ;; (offsets pinned by _Static_asserts in tests/c_abi/abi_smoke.c)
  (let [m (Memory. 112)]
    (.setInt m 0 n) (.setPointer m 8 center0-r) (.setPointer m 16 center1) (.setPointer m 24 t0t1)
    (.setPointer m 32 flags) (.setPointer m 40 mat-id)
    (.setInt m 48 n-mat) (.setPointer m 56 mat-type) (.setPointer m 64 mat-param) (.setPointer m 72 mat-tex)
    (.setInt m 80 n-tex) (.setPointer m 88 tex-type) (.setPointer m 96 tex-params) (.setPointer m 104 tex-children)
    m))

(defn- scene-ext ^Memory [{:keys [n-boundary prim-type prim-params prim-aux prim-xform n-xforms xform-ops xform-params tie-rule
                                  perlin-vectors perlin-perm n-images image-wh image-offset image-rgb]}]
  ;; struct rt_scene_ext, 120 bytes (offsets pinned by tests/c_abi/abi_smoke.c)
  (let [m (Memory. 120)]
    (.clear m)
    (.setInt m 0 120) (.setInt m 4 n-boundary)
    (.setPointer m 8 prim-type) (.setPointer m 16 prim-params) (.setPointer m 24 prim-aux) (.setPointer m 32 prim-xform)
    (.setInt m 40 n-xforms) (.setPointer m 48 xform-ops) (.setPointer m 56 xform-params)
    (.setInt m 64 tie-rule) (.setPointer m 72 perlin-vectors) (.setPointer m 80 perlin-perm)
    (.setInt m 88 n-images) (.setPointer m 96 image-wh) (.setPointer m 104 image-offset) (.setPointer m 112 image-rgb)
    m))

(defn marshal-camera [cam]
  (cond
    (instance? ThinLensCamera cam)
    [1 (floats->mem (concat (mapcat v3 [(:origin cam) (:lleft cam) (:horiz cam) (:vert cam) (:u cam) (:v cam) (:w cam)])
                            [(:aperture cam) (:t0 cam) (:t1 cam)]))]
    (instance? PinholeCamera cam)
    [0 (floats->mem (concat (mapcat v3 [(:origin cam) (:lleft cam) (:horiz cam) (:vert cam)]) (repeat 12 0)))]
    :else (throw (ex-info "unsupported camera" {:code -2}))))

(defn render
  "Drop-in for the render block core.clj:99-108: returns a byte-array of nx*ny*3 RGB, row 0 = top
  (the reference writes row ny-1-j, core.clj:105).  devices: vector of CUDA device ids."
  [camera world nx ny nr & {:keys [depth seed variant devices] :or {depth 50 seed 1 variant 1 devices [0]}}]
  (let [pctx (PointerByReference.)
        rc (.invokeInt (f "rt_create") (to-array [pctx (ints->mem devices) (int (count devices))]))
        _ (when-not (zero? rc)
            (throw (ex-info (str "rt_create: " (.invokeString (f "rt_last_error") (to-array [Pointer/NULL]) false)) {:code rc})))
        ctx (.getValue pctx)]
    (try
      (let [scene (marshal-world world)                       ; keep the Memory blocks reachable during the call
            [cam-type cam-mem] (marshal-camera camera)
            ;; page-locked result buffer (rt_host_alloc): the device writes the 8-bit image into it directly
            prgb (PointerByReference.)
            _ (check ctx (.invokeInt (f "rt_host_alloc") (to-array [ctx (long (* 3 nx ny)) prgb])) "rt_host_alloc")
            rgb (.getValue prgb)]
        (try
          (check ctx (.invokeInt (f "rt_set_scene_ex") (to-array [ctx (scene-desc scene) (scene-ext scene)])) "rt_set_scene_ex")
          (check ctx (.invokeInt (f "rt_set_camera") (to-array [ctx (int cam-type) cam-mem])) "rt_set_camera")
          (check ctx (.invokeInt (f "rt_render") (to-array [ctx (int nx) (int ny) (int nr) (int depth) (long seed)
                                                            (int variant) Pointer/NULL rgb])) "rt_render")
          (.getByteArray rgb 0 (* 3 nx ny))
          (finally (.invokeInt (f "rt_host_free") (to-array [ctx rgb])))))
      (finally (.invokeVoid (f "rt_destroy") (to-array [ctx]))))))

(defn write-ppm
  "imagez `save` has no writer for .ppm (core.clj:112); the documented CLI `lein run out.ppm …` needs one."
  [^String filename nx ny ^bytes rgb]
  (with-open [o (java.io.FileOutputStream. filename)]
    (.write o (.getBytes (str "P6\n" nx " " ny "\n255\n") "US-ASCII"))
    (.write o rgb)))
