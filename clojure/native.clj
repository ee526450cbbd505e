(ns raytrace-clj.native
  "JNA binding + marshaller for libraytrace_b200.so (include/raytrace_b200.h).

  NOT executed in the build environment (no JVM there); shipped as the source a maintainer adds to
  gonewest818/raytrace-clj next to src/raytrace_clj/core.clj.  It replaces the render block
  core.clj:99-108 with one native call and leaves arg parsing (core.clj:74-80), scene building
  (core.clj:82-90) and saving (core.clj:112) untouched.

  project.clj: add [net.java.dev.jna/jna \"5.14.0\"] to :dependencies (project.clj:6-16)."
  (:require [clojure.core.matrix :as mat])
  (:import [com.sun.jna Native Pointer Memory Structure Function NativeLibrary]
           [com.sun.jna.ptr PointerByReference]
           [raytrace_clj.hitable Sphere UVSphere MovingSphere Hitlist bvh-node]
           [raytrace_clj.shader Lambertian Metal Dielectric DiffuseLight]
           [raytrace_clj.texture Constant UVGradient Checkerboard]
           [raytrace_clj.camera ThinLensCamera PinholeCamera]))

(def ^:private lib (delay (NativeLibrary/getInstance "raytrace_b200")))
(defn- f ^Function [name] (.getFunction ^NativeLibrary @lib name))

(defn- check [ctx rc what]
  (when-not (zero? rc)
    (throw (ex-info (str what " failed: " (.invokeString (f "rt_last_error") (to-array [ctx]) false))
                    {:code rc}))))

;; ---- flatten the world: bvh-node tree (hitable.clj:97-123) -> leaves, de-duplicated by identity
(defn flatten-world [world]
  (let [seen (java.util.IdentityHashMap.) out (java.util.ArrayList.)]
    (letfn [(walk [h]
              (cond
                (instance? bvh-node h) (do (walk (:left h)) (walk (:right h)))
                (instance? Hitlist h) (run! walk (:items h))
                (or (instance? Sphere h) (instance? UVSphere h) (instance? MovingSphere h))
                (when-not (.containsKey seen h) (.put seen h true) (.add out h))
                :else (throw (ex-info (str (type h) " is outside the accelerated path (spheres only)")
                                      {:code -2}))))]
      (walk world))
    (vec out)))

(defn- floats->mem ^Memory [xs]
  (let [n (count xs) m (Memory. (max 4 (* 4 n)))]
    (.write m 0 (float-array xs) 0 n) m))
(defn- ints->mem ^Memory [xs]
  (let [n (count xs) m (Memory. (max 4 (* 4 n)))]
    (.write m 0 (int-array xs) 0 n) m))
(defn- v3 [v] [(mat/mget v 0) (mat/mget v 1) (mat/mget v 2)])

(defn marshal-world
  "world -> map of JNA Memory blocks laid out as rt_scene_desc expects"
  [world]
  (let [leaves (flatten-world world)
        texs (java.util.ArrayList.) tex-ix (java.util.IdentityHashMap.)
        mats (java.util.ArrayList.) mat-ix (java.util.IdentityHashMap.)
        add-tex (fn add-tex [t]
                  (or (.get tex-ix t)
                      (let [rec (cond
                                  (instance? Constant t) {:type 0 :p (concat (v3 (:color t)) (repeat 9 0)) :ch [-1 -1]}
                                  (instance? UVGradient t) {:type 1 :p (mapcat v3 [(:co t) (:cu t) (:cv t) (:cuv t)]) :ch [-1 -1]}
                                  (instance? Checkerboard t) {:type 2 :p (cons (:scale t) (repeat 11 0))
                                                              :ch [(add-tex (:tex0 t)) (add-tex (:tex1 t))]}
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
                                  :else (throw (ex-info (str "material " (type m) " unsupported") {:code -2})))
                            i (.size mats)]
                        (.add mats rec) (.put mat-ix m i) i)))
        rows (mapv (fn [s]
                     (if (instance? MovingSphere s)
                       {:c0r (conj (v3 (:center0 s)) (:radius s)) :c1 (conj (v3 (:center1 s)) 0)
                        :tt [(:t0 s) (:t1 s)] :flags 2 :mat (add-mat (:material s))}
                       {:c0r (conj (v3 (:center s)) (:radius s)) :c1 (conj (v3 (:center s)) 0)
                        :tt [0 1] :flags (if (instance? UVSphere s) 1 0) :mat (add-mat (:material s))}))
                   leaves)]
    {:n (count rows)
     :center0-r (floats->mem (mapcat :c0r rows)) :center1 (floats->mem (mapcat :c1 rows))
     :t0t1 (floats->mem (mapcat :tt rows)) :flags (ints->mem (map :flags rows)) :mat-id (ints->mem (map :mat rows))
     :n-mat (.size mats) :mat-type (ints->mem (map first mats)) :mat-param (floats->mem (map second mats))
     :mat-tex (ints->mem (map #(nth % 2) mats))
     :n-tex (.size texs) :tex-type (ints->mem (map :type texs)) :tex-params (floats->mem (mapcat :p texs))
     :tex-children (ints->mem (mapcat :ch texs))}))

(defn- scene-desc ^Memory [{:keys [n center0-r center1 t0t1 flags mat-id n-mat mat-type mat-param mat-tex
                                   n-tex tex-type tex-params tex-children]}]
  ;; struct rt_scene_desc on LP64: int32 + pad, 5 pointers, int32 + pad, 3 pointers, int32 + pad, 3 pointers
  (let [m (Memory. 112)]
    (.setInt m 0 n) (.setPointer m 8 center0-r) (.setPointer m 16 center1) (.setPointer m 24 t0t1)
    (.setPointer m 32 flags) (.setPointer m 40 mat-id)
    (.setInt m 48 n-mat) (.setPointer m 56 mat-type) (.setPointer m 64 mat-param) (.setPointer m 72 mat-tex)
    (.setInt m 80 n-tex) (.setPointer m 88 tex-type) (.setPointer m 96 tex-params) (.setPointer m 104 tex-children)
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
            rgb (Memory. (* 3 nx ny))]
        (check ctx (.invokeInt (f "rt_set_scene") (to-array [ctx (scene-desc scene)])) "rt_set_scene")
        (check ctx (.invokeInt (f "rt_set_camera") (to-array [ctx (int cam-type) cam-mem])) "rt_set_camera")
        (check ctx (.invokeInt (f "rt_render") (to-array [ctx (int nx) (int ny) (int nr) (int depth) (long seed)
                                                          (int variant) Pointer/NULL rgb])) "rt_render")
        (.getByteArray rgb 0 (* 3 nx ny)))
      (finally (.invokeVoid (f "rt_destroy") (to-array [ctx]))))))

(defn write-ppm
  "imagez `save` has no writer for .ppm (core.clj:112); the documented CLI `lein run out.ppm …` needs one."
  [^String filename nx ny ^bytes rgb]
  (with-open [o (java.io.FileOutputStream. filename)]
    (.write o (.getBytes (str "P6\n" nx " " ny "\n255\n") "US-ASCII"))
    (.write o rgb)))
